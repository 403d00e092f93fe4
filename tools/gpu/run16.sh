export LDAGPU_P2P_TIMEOUT_MS=10000
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_pubmed_full_1gpu_v5.json 2> gpurun_out/r02_bench_v5.err; tail -2 gpurun_out/r02_bench_v5.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_pubmed_full_1gpu_v5.json')); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['gpu_launches']); print({k:(v.get('value'),v.get('ms_per_step')) for k,v in d['secondary'].items()})"
python tools/kernel_table.py --reps 5 > gpurun_out/r02_kernel_table.json 2> gpurun_out/kt.err; tail -2 gpurun_out/kt.err; cut -c1-1500 gpurun_out/r02_kernel_table.json
