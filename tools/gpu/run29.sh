# z-step kernel code-size variants on the 400 000-document PubMed-shaped slice (z ms of the 5th sweep), then the parity suite
B="python bench.py --workload pubmed8 --docs 400000 --steps 3 --warmup 2 --no-cpu-baseline --no-secondary"
for v in "" _un "" _un; do
  echo "variant [$v]"; LDAGPU_LIBRARY=$PWD/ldagroupedgibbssampler_b200/libldagpu$v.so LDAGPU_TRACE=1 $B 2>&1 | grep "sweep 5\]" | cut -c1-120
done
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
