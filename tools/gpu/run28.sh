# validation of the tree with the split alias kernels and the full-size property checks in bench.py (driver's default command line)
export LDAGPU_P2P_TIMEOUT_MS=10000
SECONDS=0; python bench.py > gpurun_out/r02_bench_pubmed_full_1gpu_v7.json 2> gpurun_out/r02_bench_v7.err; echo "bench wall ${SECONDS}s"; grep -v "^\s" gpurun_out/r02_bench_v7.err | tail -3
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_pubmed_full_1gpu_v7.json')); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['gpu_launches']); print(d['properties']); print({k:(v.get('value'),v.get('ms_per_step'),v.get('properties'),v.get('error')) for k,v in d['secondary'].items()})"
python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-400
