python -m pytest tests -m gpu -q -x 2>&1 | tail -4
B="python bench.py --workload pubmed8 --docs 400000 --steps 3 --warmup 2 --no-cpu-baseline --no-secondary"
LDAGPU_TRACE=1 $B 2>&1 | grep "sweep 5\]"
LDAGPU_FUSE_THETA=0 LDAGPU_TRACE=1 $B 2>&1 | grep "sweep 5\]"
