# 2-GPU comparison of the two exchange modes (args: workload list)
for wl in "$@"; do
for mode in p2p nccl; do
LDAGPU_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload $wl --steps 20 --warmup 3 > gpurun_out/bench2_${wl}_$mode.json 2> gpurun_out/bench2_${wl}_$mode.err; echo "$wl $mode rc=$?"
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench2_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["config"].get("exchange"), "ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], d["timers_ms"])
    except Exception as e: print(f, "ERR", e)
PY
