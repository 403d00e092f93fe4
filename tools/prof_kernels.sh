# plain kernel table, then one full ncu capture covering every kernel of the table once
python tools/kernel_table.py --reps 5 > gpurun_out/kernel_table.json 2> gpurun_out/kernel_table.err || { tail -5 gpurun_out/kernel_table.err; exit 1; }
cat gpurun_out/kernel_table.json | python -c "import json,sys; d=json.load(sys.stdin); print('sweep', d['sweep_ms']); [print(r['phase'], r['ms'], r['achieved_gbs'], r['frac_of_peak']) for r in d['rows']]"
python tools/kernel_table.py --reps 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"counts_kernel|topic_totals|ll_doc|ll_type|lp_tokens|lp_theta|lp_phi|phi_draw|phi_segment|phi_normalise|theta_kernel|z_kernel" -s 30 -c 26 -f -o gpurun_out/prof_kernels python tools/kernel_table.py --reps 1 > gpurun_out/ncu_kernels.log 2>&1
ls -la gpurun_out/prof_kernels.ncu-rep
