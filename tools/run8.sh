# 8-GPU verification + benches (args: N)
N=${1:-8}
run() { # name, env, args...
  name=$1; shift; envs=$1; shift
  env $envs timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"
}
env LDAGPU_P2P_TIMEOUT_MS=10000 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/mg${N}_p2p.log 2>&1; echo "multigpu_check rc=$?"; grep -E "gpu_|ok" gpurun_out/mg${N}_p2p.log | cut -c1-200
run bench${N}_pubmed8_p2p LDAGPU_EXCHANGE=p2p bench.py --gpus $N --steps 20 --warmup 3
run bench${N}_pubmed8_nccl LDAGPU_EXCHANGE=nccl bench.py --gpus $N --steps 20 --warmup 3
run bench${N}_enron_p2p LDAGPU_EXCHANGE=p2p bench.py --gpus $N --workload enron --steps 20 --warmup 3
run bench${N}_enron_nccl LDAGPU_EXCHANGE=nccl bench.py --gpus $N --workload enron --steps 20 --warmup 3
run bench${N}_wiki8_p2p LDAGPU_EXCHANGE=p2p bench.py --gpus $N --workload wiki8 --steps 5 --warmup 3
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench[0-9]_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["config"].get("exchange"), "n", d["n_gpus"], "ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], {k:round(v,1) for k,v in d["timers_ms"].items()})
    except Exception as e: print(f, "ERR", e)
PY
