# parity tests + the three dense benches on one GPU
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pt.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pt.log
for wl in "$@"; do
python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/q_$wl.json 2> gpurun_out/q_$wl.err; echo "$wl rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/q_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        n=d["steps"]+d["warmup"]+min(d["warmup"],2)+d["steps"]
        print(f, "ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], "zk ms", round(d["roofline"]["kernel_ms_per_launch"],3), {k:round(v/n,3) for k,v in d["timers_ms"].items()})
    except Exception as e: print(f, "ERR", e)
PY
