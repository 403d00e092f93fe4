B="python bench.py --workload wiki8 --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
$B > gpurun_out/plain_wiki8.log 2> gpurun_out/plain_wiki8.err && ncu --set full --clock-control none --import-source on -k regex:alias_pair -s 3 -c 1 -f -o gpurun_out/r02_ncu_alias_wiki8_v5 $B > gpurun_out/ncu_alias.log 2>&1
tail -1 gpurun_out/ncu_alias.log | cut -c1-300
