python tools/kernel_table.py --reps 5 > gpurun_out/r02_kernel_table.json 2> gpurun_out/kt.err; tail -2 gpurun_out/kt.err
python -c "
import json; d=json.load(open('gpurun_out/r02_kernel_table.json')); print(d['sweep_ms']); print([(r['phase'], r['ms']) for r in d['rows']])"
