# final kernels: suite, smoke, ncu --set full of the z-step kernel and the launch list on the 400 000-document slice, default bench
python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
B="python bench.py --workload pubmed8 --docs 400000 --steps 2 --warmup 2 --no-cpu-baseline --no-secondary"
$B > gpurun_out/plain_pubmed8.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:z_kernel -s 3 -c 1 -f -o gpurun_out/r02_ncu_z_pubmed8_final2 $B > gpurun_out/ncu_pubmed8.log 2>&1
tail -1 gpurun_out/ncu_pubmed8.log | cut -c1-200
B="python bench.py --workload pubmed8 --docs 400000 --steps 2 --warmup 1 --no-cpu-baseline --no-secondary"
$B > gpurun_out/plain_l.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_pubmed8.csv $B > gpurun_out/ncu_l.log 2>&1
tail -1 gpurun_out/ncu_l.log | cut -c1-200
python bench.py > gpurun_out/r02_bench_pubmed_full_1gpu_v8.json 2> gpurun_out/r02_bench_v8.err; grep -v "^\s" gpurun_out/r02_bench_v8.err | tail -2
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_pubmed_full_1gpu_v8.json')); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['gpu_launches']); print({k:(v.get('value'),v.get('ms_per_step'),v.get('error')) for k,v in d['secondary'].items()})"
