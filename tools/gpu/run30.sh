# same box: instruction-footprint variants of the z-step kernel (orig = 4x unrolled token loop, two drain sites, two request sites)
B="python bench.py --workload pubmed8 --docs 400000 --steps 3 --warmup 2 --no-cpu-baseline --no-secondary"
for v in _orig _nu _nu_sd _nu_mr "" _orig _nu _nu_sd _nu_mr ""; do
  echo "variant [$v] $(LDAGPU_LIBRARY=$PWD/ldagroupedgibbssampler_b200/libldagpu$v.so LDAGPU_TRACE=1 $B 2>&1 | grep 'sweep 5\]' | cut -c24-60)"
done
