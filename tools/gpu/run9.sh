export LDAGPU_P2P_TIMEOUT_MS=10000
python -m pytest tests -m gpu -q 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_pubmed_full_1gpu_v3.json 2> gpurun_out/r02_bench_v3.err; tail -2 gpurun_out/r02_bench_v3.err; cut -c1-300 gpurun_out/r02_bench_pubmed_full_1gpu_v3.json
for wl in pubmed8 nips enron; do
  extra=""; [ $wl = pubmed8 ] && extra="--docs 400000"
  B="python bench.py --workload $wl $extra --steps 2 --warmup 2 --no-cpu-baseline --no-secondary"
  $B > gpurun_out/plain_$wl.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:z_kernel -s 3 -c 1 -o gpurun_out/r02_ncu_z_$wl $B > gpurun_out/ncu_$wl.log 2>&1
  tail -1 gpurun_out/ncu_$wl.log
done
B="python bench.py --workload pubmed8 --docs 400000 --steps 2 --warmup 1 --no-cpu-baseline --no-secondary"
$B > gpurun_out/plain_l.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_pubmed8.csv $B > gpurun_out/ncu_l.log 2>&1
tail -1 gpurun_out/ncu_l.log
