# alias-table build rewritten (stack-ordered scratch, queued fetches): sparse parity tests, then the phase times per variant
timeout 500 python -m pytest tests -m gpu -x -q -k "sparse or spalias or alias or polya or product" 2>&1 | tail -4
B="python bench.py --workload wiki8 --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
for v in "" _c6 _c2; do
  echo "variant [$v]"; LDAGPU_LIBRARY=$PWD/ldagroupedgibbssampler_b200/libldagpu$v.so LDAGPU_TRACE=1 $B 2>&1 | grep "sweep 5\]" | cut -c1-260
done
