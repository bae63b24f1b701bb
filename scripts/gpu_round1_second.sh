mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --tb=short -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python scripts/accuracy_probe.py > gpurun_out/accuracy_fast.log 2>&1
NOC_LIB=$PWD/neuraloc_b200/libnoc_b200_precise.so timeout 600 python scripts/accuracy_probe.py > gpurun_out/accuracy_precise.log 2>&1
NOC_LIB=$PWD/neuraloc_b200/libnoc_b200_precise.so timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_swap12_precise.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_swap12.csv python bench.py --steps 2 --warmup 3 --samples 262144 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:rollout_kernel -s 1 -c 1 -f -o gpurun_out/prof_swap12 python bench.py --steps 1 --warmup 3 --samples 131072 --no-cpu-baseline > gpurun_out/ncu_swap12.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:rollout_kernel -s 1 -c 1 -f -o gpurun_out/prof_swarm50 python bench.py --steps 1 --warmup 3 --workload swarm50 --samples 16384 --no-cpu-baseline > gpurun_out/ncu_swarm50.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:rollout_kernel -s 1 -c 1 -f -o gpurun_out/prof_quad python bench.py --steps 1 --warmup 3 --workload singlequad --samples 131072 --no-cpu-baseline > gpurun_out/ncu_quad.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/accuracy_fast.log; ls -la gpurun_out
