N=8
export LDAGPU_P2P_TIMEOUT_MS=10000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_pubmed_${N}gpu_strong_v3.json 2> gpurun_out/b${N}.err; tail -2 gpurun_out/b${N}.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_pubmed_${N}gpu_strong_v3.json')); print('N=$N', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['run']['exchange'], d['clocks'])"
python bench.py --gpus $N --impl reference --steps 2 --warmup 1 | cut -c1-300
