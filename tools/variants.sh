# usage: variants.sh <workload> <tag>...   (libldagpu_<tag>.so beside the default; "base" = default)
wl=$1; shift
for tag in "$@"; do
  lib=$PWD/ldagroupedgibbssampler_b200/libldagpu_$tag.so; [ "$tag" = base ] && lib=$PWD/ldagroupedgibbssampler_b200/libldagpu.so
  LDAGPU_LIBRARY=$lib python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/var_${wl}_$tag.json 2> gpurun_out/var_${wl}_$tag.err
  python - "$tag" gpurun_out/var_${wl}_$tag.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
n=d["steps"]+d["warmup"]+min(d["warmup"],2)+d["steps"]
print(sys.argv[1], "ms/step", round(d["ms_per_step"],3), "zk ms", round(d["roofline"]["kernel_ms_per_launch"],3), {k:round(v/n,3) for k,v in d["timers_ms"].items()})
PY
done
