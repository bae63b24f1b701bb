# accumulate-bias compensation of the m <= 128 tensor kernel: mean-cost errors against the fp64 oracle for several factors,
# step counts and sample sets (NOC_TC_BIAS multiplies the compensation)
for cfg in "30 5" "80 11"; do set -- $cfg
for w in singlequad swap12 softcorridor swap2; do
for f in 0.5 1 1.5 2; do echo "== $w nt=$1 seed=$2 BIAS=$f"; NOC_TC_BIAS=$f PROBE_NT=$1 PROBE_SEED=$2 PROBE_N=1024 PROBE_ONLY=$w timeout 300 python scripts/accuracy_probe.py 2>&1 | grep -v library | head -2; done; done; done
