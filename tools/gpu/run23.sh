# alias-table build of the Wikipedia-shaped K=10 000 shard: one ncu --set full capture with source counters
B="python bench.py --workload wiki8 --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
LDAGPU_TRACE=1 $B > gpurun_out/plain_wiki8.log 2> gpurun_out/plain_wiki8.err && ncu --set full --clock-control none --import-source on -k regex:alias_build -s 3 -c 1 -o gpurun_out/r02_ncu_alias_wiki8 $B > gpurun_out/ncu_alias.log 2>&1
tail -1 gpurun_out/ncu_alias.log | cut -c1-300
tail -3 gpurun_out/plain_wiki8.err | cut -c1-400
