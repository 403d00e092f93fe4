python -m pytest tests -m gpu -q -x 2>&1 | tail -8
B="python bench.py --workload pubmed8 --docs 400000 --steps 3 --warmup 2 --no-cpu-baseline"
for v in "" _w7m3 _w10m2 _w6m3 _w8m2; do
  echo "variant [$v]"; LDAGPU_LIBRARY=$PWD/ldagroupedgibbssampler_b200/libldagpu$v.so LDAGPU_TRACE=1 $B 2>&1 | grep "sweep 5\]" 
done
