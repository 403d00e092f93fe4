B="python bench.py --workload pubmed8 --docs 400000 --steps 3 --warmup 2 --no-cpu-baseline --no-secondary"
for v in "" _w8m4 _w7m4 _w9m3 _w10m3; do
  echo "variant [$v]"; LDAGPU_LIBRARY=$PWD/ldagroupedgibbssampler_b200/libldagpu$v.so LDAGPU_TRACE=1 $B 2>&1 | grep "sweep 5\]" 
done
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_pubmed_v2.json 2> gpurun_out/r02_bench_pubmed_v2.err; tail -3 gpurun_out/r02_bench_pubmed_v2.err; cut -c1-1500 gpurun_out/r02_bench_pubmed_v2.json
python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-600
