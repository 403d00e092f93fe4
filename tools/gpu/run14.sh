export LDAGPU_P2P_TIMEOUT_MS=10000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
LDAGPU_STRESS_DUMP_AFTER=60 LDAGPU_EXCHANGE=nccl timeout 200 $TR --master-port 29553 tests/multigpu_check.py --stress 60 2>&1 | grep -v "^frame\|Warning\|warn" | tail -60 | cut -c1-250
