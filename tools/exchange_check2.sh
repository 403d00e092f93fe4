set -x
for mode in p2p nccl; do
LDAGPU_EXCHANGE=$mode LDAGPU_P2P_TIMEOUT_MS=8000 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/mg2_$mode.log 2>&1; echo rc=$? >> gpurun_out/mg2_$mode.log
done
tail -n 8 gpurun_out/mg2_p2p.log gpurun_out/mg2_nccl.log
