export LDAGPU_P2P_TIMEOUT_MS=10000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
LDAGPU_STRESS_DUMP_AFTER=150 LDAGPU_EXCHANGE=nccl timeout 260 $TR --master-port 29553 tests/multigpu_check.py --stress 200 2>&1 | grep -v "^frame\|Warning\|warn\|^\*\*\*\|OMP_NUM" | tail -30 | cut -c1-250 | tee gpurun_out/r02_multigpu_stress_2gpu_nccl.log
LDAGPU_STRESS_DUMP_AFTER=150 timeout 260 $TR --master-port 29555 tests/multigpu_check.py --stress 200 2>&1 | grep -v "^frame\|Warning\|warn\|^\*\*\*\|OMP_NUM" | tail -30 | cut -c1-250 | tee gpurun_out/r02_multigpu_stress_2gpu_p2p.log
