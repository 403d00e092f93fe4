nvidia-smi -L | head -3
python tests/singleproc_multigpu_check.py --gpus 2 --stress 40 2>&1 | tail -12
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/b2.err | cut -c1-1400; tail -2 gpurun_out/b2.err
