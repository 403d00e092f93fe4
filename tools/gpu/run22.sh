B="python bench.py --workload pubmed8 --docs 400000 --steps 2 --warmup 2 --no-cpu-baseline --no-secondary"
$B > gpurun_out/plain_pubmed8.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:z_kernel -s 3 -c 1 -o gpurun_out/r02_ncu_z_pubmed8_final $B > gpurun_out/ncu_pubmed8.log 2>&1
tail -1 gpurun_out/ncu_pubmed8.log
B="python bench.py --workload pubmed8 --docs 400000 --steps 2 --warmup 1 --no-cpu-baseline --no-secondary"
$B > gpurun_out/plain_l.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_pubmed8.csv $B > gpurun_out/ncu_l.log 2>&1
tail -1 gpurun_out/ncu_l.log
