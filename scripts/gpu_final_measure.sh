# Round-end evidence: default bench line (both arms), launch list, ncu --set full of the three rollout kernels.
TAG=${1:-final}
mkdir -p gpurun_out
python bench.py > gpurun_out/BENCH_default_$TAG.json 2> gpurun_out/BENCH_default_$TAG.err
python bench.py --impl reference --steps 1 > gpurun_out/BENCH_reference_$TAG.json 2>> gpurun_out/BENCH_default_$TAG.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/clocks_after_$TAG.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_default_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
bash scripts/gpu_prof.sh $TAG swap12 131072
bash scripts/gpu_prof.sh $TAG swarm50 16384
bash scripts/gpu_prof.sh $TAG singlequad 131072
timeout 600 ncu --set full --import-source on --clock-control none -k regex:rollout_vec_kernel -s 2 -c 1 -f -o gpurun_out/prof_vec_swap12_$TAG python bench.py --latency --workload swap12 > gpurun_out/ncu_vec_$TAG.log 2>&1
python scripts/peaks.py > gpurun_out/peaks_$TAG.txt 2>&1
cat gpurun_out/BENCH_default_$TAG.json | cut -c1-1500; cat gpurun_out/BENCH_reference_$TAG.json | cut -c1-600; cat gpurun_out/peaks_$TAG.txt
