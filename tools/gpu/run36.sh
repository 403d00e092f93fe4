python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
LDAGPU_TRACE=1 python bench.py --workload wiki8 --steps 3 --warmup 3 --no-cpu-baseline --no-secondary 2>&1 | grep "sweep 6\]" | cut -c1-260
