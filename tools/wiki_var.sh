for tag in "$@"; do
  lib=$PWD/ldagroupedgibbssampler_b200/libldagpu_$tag.so; [ "$tag" = base ] && lib=$PWD/ldagroupedgibbssampler_b200/libldagpu.so
  LDAGPU_LIBRARY=$lib LDAGPU_TRACE=1 python bench.py --workload wiki8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/wv_$tag.json 2> gpurun_out/wv_$tag.err
  echo $tag; grep "ldagpu rank" gpurun_out/wv_$tag.err | tail -1
done
