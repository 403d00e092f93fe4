# one --set full capture of the sparse z kernel in steady state (9th sweep) on the full Wikipedia-shaped shard
cmd="python bench.py --workload wiki8 --steps 2 --warmup 8 --no-cpu-baseline"
$cmd > gpurun_out/plain_wiki8s.log 2> gpurun_out/plain_wiki8s.err || { tail -3 gpurun_out/plain_wiki8s.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:z_spalias -s 8 -c 1 -f -o gpurun_out/prof_wiki8s $cmd > gpurun_out/ncu_f_wiki8s.log 2>&1
ls -la gpurun_out/prof_wiki8s.ncu-rep
