# multi-GPU verification + benches: runN.sh <N> <workload:steps>...
N=$1; shift
env LDAGPU_P2P_TIMEOUT_MS=15000 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/mg${N}.log 2>&1; echo "multigpu_check rc=$?"; grep -E "gpu_|ok" gpurun_out/mg${N}.log | cut -c1-120
for spec in "$@"; do
  wl=${spec%%:*}; st=${spec##*:}
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload $wl --steps $st --warmup 3 > gpurun_out/final${N}_$wl.json 2> gpurun_out/final${N}_$wl.err; echo "$wl rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/final[0-9]_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["config"].get("exchange"), "n", d["n_gpus"], "ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
