export LDAGPU_P2P_TIMEOUT_MS=10000
B="python bench.py --workload pubmed8 --docs 400000 --steps 3 --warmup 2 --no-cpu-baseline --no-secondary"
for v in "" _w4 _w1; do
  echo "variant [$v]"; LDAGPU_LIBRARY=$PWD/ldagroupedgibbssampler_b200/libldagpu$v.so LDAGPU_TRACE=1 $B 2>&1 | grep "sweep 5\]" 
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_pubmed_2gpu_strong.json 2> gpurun_out/b2.err; tail -1 gpurun_out/b2.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_pubmed_2gpu_strong.json')); print('N=2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
$TR --master-port 29551 tests/multigpu_check.py --stress 200 2>&1 | grep -E "stress|Error|error" | tee gpurun_out/r02_multigpu_stress_2gpu_p2p.log
LDAGPU_EXCHANGE=nccl $TR --master-port 29553 tests/multigpu_check.py --stress 200 2>&1 | grep -E "stress|Error|error" | tee gpurun_out/r02_multigpu_stress_2gpu_nccl.log
