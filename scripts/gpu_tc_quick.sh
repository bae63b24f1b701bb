# quick loop for the m <= 128 tensor kernel: parity tests, accuracy against the fp64 oracle, throughput
timeout 900 python -m pytest tests/test_gpu_tc_rollout.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4
for w in singlequad swap12 softcorridor swap2; do PROBE_ONLY=$w timeout 300 python scripts/accuracy_probe.py 2>&1 | grep -v library | head -2; done
for w in swap12 singlequad softcorridor swap2; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d[\"config\"][\"workload\"][:12], \"%.4g\" % d[\"value\"], \"e2e %.4g\" % d[\"e2e\"][\"value\"], d[\"clocks\"][\"sm_mhz\"], d[\"clocks\"][\"power_w_max\"])"; done
