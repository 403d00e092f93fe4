# launch lists + one full capture per workload (1 GPU).  usage: prof_small.sh <tag> <workload> <kernel regex> [extra bench args]
tag=$1; wl=$2; kre=$3; shift 3
cmd="python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline $*"
$cmd > gpurun_out/plain_$tag.log 2> gpurun_out/plain_$tag.err || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $cmd > gpurun_out/ncu_l_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$kre -s 2 -c 2 -f -o gpurun_out/prof_$tag $cmd > gpurun_out/ncu_f_$tag.log 2>&1
ls -la gpurun_out/prof_$tag.ncu-rep
