# usage: bash scripts/gpu_bench_all.sh <tag>   -- parity tests + smoke + one bench line per workload into gpurun_out/
TAG=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log
for W in swap12 softcorridor swap2; do timeout 300 python bench.py --steps 3 --warmup 3 --workload $W --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_${W}_$TAG.json; done
timeout 300 python bench.py --steps 2 --warmup 3 --workload singlequad --samples 524288 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_singlequad_$TAG.json
timeout 300 python bench.py --steps 2 --warmup 3 --workload swarm50 --samples 65536 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_swarm50_$TAG.json
timeout 300 python bench.py --steps 2 --warmup 3 --workload config5 --samples 16384 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_config5_$TAG.json
tail -3 gpurun_out/pytest_gpu_$TAG.log; tail -2 gpurun_out/smoke_$TAG.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*_$TAG.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print("%-40s %.4e sample-steps/s  %.1f ms/step  %.2f TF (%.1f%% of %.1f)  e2e %.4e" % (f.split("/")[-1], d["value"], d["ms_per_step"], r["achieved"], 100*r["frac"], r["peak"], d["e2e"]["value"]))
    except Exception as e: print(f, "ERR", e, open(f).read()[-300:])
PY
