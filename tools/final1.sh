# final single-GPU campaign: tests, smoke, benches of every workload, reference arm, ncu evidence
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_pt.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_pt.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/final_pubmed8.json 2> gpurun_out/final_pubmed8.err; echo "pubmed8 rc=$?"
python bench.py --workload nips > gpurun_out/final_nips.json 2> gpurun_out/final_nips.err; echo "nips rc=$?"
python bench.py --workload enron > gpurun_out/final_enron.json 2> gpurun_out/final_enron.err; echo "enron rc=$?"
python bench.py --workload wiki8 --steps 8 --warmup 3 > gpurun_out/final_wiki8.json 2> gpurun_out/final_wiki8.err; echo "wiki8 rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err; echo "reference rc=$?"
cmd="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$cmd > gpurun_out/final_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv $cmd > gpurun_out/final_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"z_kernel|theta_kernel|phi_draw" -s 3 -c 6 -f -o gpurun_out/final_prof $cmd > gpurun_out/final_ncu_f.log 2>&1
ls -la gpurun_out/final_prof.ncu-rep
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/final_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], "cpu %.3e"%d.get("cpu_baseline",{}).get("value",0), d.get("clocks"))
    except Exception as e: print(f, "ERR", e)
PY
