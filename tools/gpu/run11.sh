N=8
export LDAGPU_P2P_TIMEOUT_MS=10000
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_pubmed_${N}gpu_strong.json 2> gpurun_out/b${N}.err; tail -2 gpurun_out/b${N}.err | cut -c1-300
$TR --master-port 29547 bench.py --gpus $N --steps 50 --warmup 5 --workload enron > gpurun_out/r02_bench_enron_${N}gpu_strong.json 2>> gpurun_out/b${N}.err
$TR --master-port 29549 bench.py --gpus $N --steps 5 --warmup 3 --workload wiki8 > gpurun_out/r02_bench_wiki8_${N}gpu_weak.json 2>> gpurun_out/b${N}.err
for f in pubmed_${N}gpu_strong enron_${N}gpu_strong wiki8_${N}gpu_weak; do python -c "
import json; d=json.load(open('gpurun_out/r02_bench_$f.json')); print('$f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['run']['exchange'], d['clocks']['reasons'])"; done
python tests/singleproc_multigpu_check.py --gpus $N --stress 60 2>&1 | tail -5 | cut -c1-400 | tee gpurun_out/r02_singleproc_multigpu_${N}gpu.log
$TR --master-port 29543 tests/multigpu_check.py 2>&1 | grep -v Warning | tail -6 | cut -c1-400 | tee gpurun_out/r02_multigpu_check_${N}gpu.log
