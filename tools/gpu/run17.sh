export LDAGPU_P2P_TIMEOUT_MS=10000
python -m pytest tests/test_gpu_parity.py tests/test_gpu_product_paths.py tests/test_gpu_streamed_readback.py -m gpu -q -x 2>&1 | tail -3
B="python bench.py --workload pubmed8 --docs 400000 --steps 3 --warmup 2 --no-cpu-baseline --no-secondary"
for v in "" _t4 _t5 _t7; do
  echo "variant [$v]"; LDAGPU_LIBRARY=$PWD/ldagroupedgibbssampler_b200/libldagpu$v.so LDAGPU_TRACE=1 $B 2>&1 | grep "sweep 5\]" 
done
