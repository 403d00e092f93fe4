# usage: run10.sh N
N=$1
export LDAGPU_P2P_TIMEOUT_MS=10000
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_pubmed_${N}gpu_strong.json 2> gpurun_out/b${N}.err; tail -2 gpurun_out/b${N}.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_pubmed_${N}gpu_strong.json')); print('N=$N', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['run']['exchange'], d['clocks'])"
python tests/singleproc_multigpu_check.py --gpus $N --stress 60 2>&1 | tail -5 | cut -c1-400 | tee gpurun_out/r02_singleproc_multigpu_${N}gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 tests/multigpu_check.py 2>&1 | grep -v Warning | tail -6 | cut -c1-400 | tee gpurun_out/r02_multigpu_check_${N}gpu.log
if [ $N -le 4 ]; then
python -m pytest tests/test_gpu_multirank.py tests/test_outputs.py -m gpu -q 2>&1 | tail -4
LDAGPU_EXCHANGE=nccl python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --steps 10 --warmup 3 --no-secondary 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('nccl N=$N', d['value'], d['ms_per_step'], d['run']['exchange'])"
fi
