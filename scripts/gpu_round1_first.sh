mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep "Model name" >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_swap12.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_swap12.log
timeout 300 python bench.py --steps 2 --warmup 3 --workload singlequad --samples 262144 --no-cpu-baseline > gpurun_out/bench_quad.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --workload swarm50 --samples 65536 --no-cpu-baseline > gpurun_out/bench_swarm.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --workload softcorridor --no-cpu-baseline > gpurun_out/bench_soft.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -2 gpurun_out/bench_swap12.log
