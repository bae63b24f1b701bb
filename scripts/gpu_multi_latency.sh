# N=2 NCCL run of the default bench + batch-1 latency of every problem (GPU 0)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_swap12_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --workload swarm50 --samples 65536 2>&1 | tail -1 > gpurun_out/bench_swarm50_n2.json
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_swap12_n1.json
for W in softcorridor swap2 swap12 singlequad swarm50; do timeout 300 python bench.py --latency --workload $W 2>&1 | tail -1 > gpurun_out/latency_$W.json; done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*_n?.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]; print(f.split("/")[-1], "N=%d"%d["n_gpus"], "%.4e"%d["value"], "%.1f ms"%d["ms_per_step"], "%.1f%%"%(100*r["frac"]), "e2e %.4e"%d["e2e"]["value"], d.get("clocks"))
    except Exception as e: print(f, "ERR", e, open(f).read()[-400:])
for f in sorted(glob.glob("gpurun_out/latency_*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], "host wall %.3f ms (p10 %.3f, p90 %.3f), device %.3f ms, cpu 1-thread %.1f ms, Jc %.6e vs %.6e" % (d["value"], d["host_wall_ms"]["p10"], d["host_wall_ms"]["p90"], d["device_ms_median"], d["cpu_baseline"]["value"], d["Jc"], d["Jc_cpu"]))
    except Exception as e: print(f, "ERR", e, open(f).read()[-400:])
PY
