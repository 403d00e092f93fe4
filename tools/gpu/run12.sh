export LDAGPU_P2P_TIMEOUT_MS=10000
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_pubmed_full_1gpu_v4.json 2> gpurun_out/r02_bench_v4.err; tail -2 gpurun_out/r02_bench_v4.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_pubmed_full_1gpu_v4.json')); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print({k:d['roofline'][k] for k in ('achieved','frac','traffic','achieved_model','frac_model','fetches_per_token','l2_to_sm_gbs','kernel_ms_per_launch','kernel_share_of_step')}); print(d['roofline']['binding']); print({k:(v.get('value'),v.get('ms_per_step')) for k,v in d['secondary'].items()}); print(d['timers_ms'], d['gpu_launches'])"
python bench.py --impl reference --steps 3 --warmup 1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ref', d['value'], d['ms_per_step'], d['cpu_baseline']['cores'], d['config']==json.load(open('gpurun_out/r02_bench_pubmed_full_1gpu_v4.json'))['config'])"
