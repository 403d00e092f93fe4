# alias build split into classify + pair kernels: whole GPU suite, smoke, wiki8 phase times for both Phi samplers
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for w in wiki8 wiki8_polya; do
  LDAGPU_TRACE=1 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_${w}_1gpu_v2.json 2> gpurun_out/tr_$w.err
  grep "sweep 7\]" gpurun_out/tr_$w.err | cut -c1-260
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_${w}_1gpu_v2.json')); print('$w', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
done
