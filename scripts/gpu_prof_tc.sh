# usage: bash scripts/gpu_prof_tc.sh <tag> <workload> <n> [lib]  -- one ncu --set full capture of the tensor-core rollout kernel
TAG=$1; W=$2; N=$3; LIBF=${4:-libnoc_b200.so}
mkdir -p gpurun_out
NOC_LIB=/root/repo/neuraloc_b200/$LIBF NOC_TC=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:rollout_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_tc_${W}_$TAG python bench.py --steps 1 --warmup 3 --workload $W --samples $N --no-cpu-baseline > gpurun_out/ncu_tc_${W}_$TAG.log 2>&1
tail -2 gpurun_out/ncu_tc_${W}_$TAG.log | cut -c1-300
