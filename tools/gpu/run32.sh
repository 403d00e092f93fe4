# same box: register / occupancy variants of the z-step kernel after the token loop lost its unrolling
B="python bench.py --workload pubmed8 --docs 400000 --steps 3 --warmup 2 --no-cpu-baseline --no-secondary"
for v in "" _t2 _t4 _w12 _w6m4 ""; do
  echo "variant [$v] $(LDAGPU_LIBRARY=$PWD/ldagroupedgibbssampler_b200/libldagpu$v.so LDAGPU_TRACE=1 $B 2>&1 | grep 'sweep 5\]' | cut -c24-60)"
done
