# Short round-end refresh after tensor-kernel changes: default bench (both arms), launch list, ncu --set full of the two
# tensor-core rollout kernels, one bench line per workload.  (FMA / small-batch kernel captures: scripts/gpu_final_measure.sh.)
TAG=${1:-final}
mkdir -p gpurun_out
python bench.py > gpurun_out/BENCH_default_$TAG.json 2> gpurun_out/BENCH_default_$TAG.err
python bench.py --impl reference --steps 1 > gpurun_out/BENCH_reference_$TAG.json 2>> gpurun_out/BENCH_default_$TAG.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/clocks_after_$TAG.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_default_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
bash scripts/gpu_prof.sh ${TAG}_tc swap12 262144 rollout_tc_kernel 1
bash scripts/gpu_prof.sh ${TAG}_tc singlequad 262144 rollout_tc_kernel 1
for W in softcorridor swap2 singlequad; do timeout 300 python bench.py --steps 3 --warmup 3 --workload $W --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_${W}_$TAG.json; done
timeout 300 python bench.py --steps 2 --warmup 3 --workload swarm50 --samples 65536 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_swarm50_$TAG.json
cat gpurun_out/BENCH_default_$TAG.json | cut -c1-1800; cat gpurun_out/BENCH_reference_$TAG.json | cut -c1-400
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*_$TAG.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]; print(f.split("/")[-1], "%.4e"%d["value"], "%.1f ms"%d["ms_per_step"], r["bound"], "%.1f TF %.1f%%"%(r["achieved"],100*r["frac"]), "e2e %.4e"%d["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
