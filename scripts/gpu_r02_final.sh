# Round-2 final measurement batch on one B200 (after the training / baseline kernels and the fp16 x 2 split of the m <= 128 kernel).
mkdir -p gpurun_out
T=${1:-r2f}
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 2000 > gpurun_out/clocks_$T.csv &
SMI=$!
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/BENCH_default_$T.json 2> gpurun_out/BENCH_default_$T.err
tail -c 300 gpurun_out/BENCH_default_$T.err
kill $SMI
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/BENCH_reference_$T.json 2>/dev/null
for w in softcorridor swap2 config5; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_$T.json 2>/dev/null
done
for w in softcorridor swap2 swap12 singlequad swarm50; do
  timeout 300 python bench.py --latency --workload $w > gpurun_out/latency_${w}_$T.json 2>/dev/null
  timeout 400 python bench.py --train --workload $w --steps 5 --warmup 3 > gpurun_out/train_${w}_$T.json 2>/dev/null
done
timeout 600 python bench.py --intermediates --workload swarm50 --steps 3 --warmup 1 > gpurun_out/bench_intermediates_swarm50_$T.json 2>/dev/null
timeout 600 python bench.py --intermediates --workload swap12 --steps 3 --warmup 1 > gpurun_out/bench_intermediates_swap12_$T.json 2>/dev/null
# launch list of a short default bench (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default_$T.csv python bench.py --steps 2 --warmup 1 --samples 65536 --no-extra --no-cpu-baseline > gpurun_out/ncu_launches_$T.log 2>&1
# one full capture per kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_ts -c 1 -f -o gpurun_out/prof_ts_swarm50_$T python scripts/ts_run.py 9472 20 > gpurun_out/ncu_ts_$T.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_tc_swap12_$T python bench.py --steps 1 --warmup 3 --workload swap12 --samples 262144 --no-cpu-baseline > gpurun_out/ncu_tc_swap12_$T.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_tc_singlequad_$T python bench.py --steps 1 --warmup 3 --workload singlequad --samples 262144 --no-cpu-baseline > gpurun_out/ncu_tc_singlequad_$T.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_grad_kernel -s 1 -c 1 -f -o gpurun_out/prof_grad_swarm50_$T python bench.py --train --workload swarm50 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_grad_$T.log 2>&1
# compute-sanitizer over the two new kernels
for tool in memcheck racecheck synccheck; do
  for c in "softcorridor 13 3 f32" "swap12 13 2 f64" "singlequad 13 2 f32" "swarm50 11 1 f32"; do
    tag=$(echo $c | cut -d' ' -f1)
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_train.py $c > gpurun_out/sanitizer_${tool}_train_${tag}.log 2>&1
    echo "== $tool train_$tag: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_train_${tag}.log | tail -1) | $(grep sanitize_train gpurun_out/sanitizer_${tool}_train_${tag}.log | tail -1)"
  done
done
ls -la gpurun_out/*_$T* | awk '{print $5, $9}'
