N=$1; wl=$2; shift 2
for mode in p2p nccl; do
LDAGPU_TRACE=1 LDAGPU_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload $wl --steps 3 --warmup 2 "$@" > gpurun_out/trace_${wl}_$mode.json 2> gpurun_out/trace_${wl}_$mode.err; echo "$wl $mode rc=$?"
grep "ldagpu rank" gpurun_out/trace_${wl}_$mode.err | sort -k5n -k3n | sed -n 1,200p | awk '$5>=4 && $5<=5' 
done
