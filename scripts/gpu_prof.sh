# usage: bash scripts/gpu_prof.sh <tag> <workload> <n> [kernel-regex] [NOC_TC]  -- one ncu --set full capture of a rollout kernel
TAG=$1; W=$2; N=$3; K=${4:-rollout_kernel}; TC=${5:-1}
mkdir -p gpurun_out
NOC_TC=$TC timeout 900 ncu --set full --import-source on --clock-control none -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_${W}_$TAG python bench.py --steps 1 --warmup 3 --workload $W --samples $N --no-cpu-baseline > gpurun_out/ncu_${W}_$TAG.log 2>&1
tail -2 gpurun_out/ncu_${W}_$TAG.log | cut -c1-300
