# two GPUs: the multi-GPU paths with the split alias kernels (single process and one process per GPU), 2-GPU pytest cases
export LDAGPU_P2P_TIMEOUT_MS=10000
timeout 300 python tests/singleproc_multigpu_check.py --gpus 2 --stress 20 > gpurun_out/r02_singleproc_multigpu_2gpu_v2.log 2>&1; tail -8 gpurun_out/r02_singleproc_multigpu_2gpu_v2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tests/multigpu_check.py > gpurun_out/r02_multigpu_check_2gpu_v2.log 2>&1; tail -6 gpurun_out/r02_multigpu_check_2gpu_v2.log
timeout 400 python -m pytest tests -m gpu -q -x -k "multi or rank or shard" 2>&1 | tail -3
