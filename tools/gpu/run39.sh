# Wikipedia-shaped sparse K=10 000 on 8 GPUs (weak: 1 B tokens), split alias-table kernels
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --workload wiki8 --steps 5 --warmup 3 > gpurun_out/r02_bench_wiki8_8gpu_weak_v2.json 2> gpurun_out/w8.err; tail -2 gpurun_out/w8.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_wiki8_8gpu_weak_v2.json')); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['run']['exchange'], d['clocks'], d['properties'])"
