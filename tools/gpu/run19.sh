export LDAGPU_P2P_TIMEOUT_MS=10000
for wl in wiki8 wiki8_polya; do
LDAGPU_TRACE=1 python bench.py --workload $wl --steps 4 --warmup 4 --no-cpu-baseline --no-secondary 2> gpurun_out/tr_$wl.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$wl', d['value'], d['ms_per_step'], d['roofline']['mean_nnz_d'], d['roofline']['kernel_ms_per_launch'])"
grep "sweep 8\]" gpurun_out/tr_$wl.err | cut -c1-300
done
