# the driver's 4-GPU command line on the final tree (properties block under world > 1)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_bench_pubmed_4gpu_strong_v2.json 2> gpurun_out/b4.err; tail -3 gpurun_out/b4.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_pubmed_4gpu_strong_v2.json')); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['run']['exchange'], d['properties'])"
