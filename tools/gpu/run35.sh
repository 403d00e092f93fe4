# pairing kernel reading both stacks in aligned groups of four, low stack one group ahead: sparse parity tests + phase times
timeout 500 python -m pytest tests -m gpu -x -q -k "sparse or spalias or alias or polya or product" 2>&1 | tail -3
LDAGPU_TRACE=1 python bench.py --workload wiki8 --steps 3 --warmup 3 --no-cpu-baseline --no-secondary 2>&1 | grep "sweep 6\]" | cut -c1-260
