# usage: bash scripts/gpu_quick.sh "<pytest -k expr>" W:N [W:N ...]  -- subset of parity tests + short benches
K="$1"; shift
python -m pytest tests -m gpu -q -x -k "$K" -p no:cacheprovider 2>&1 | tail -3
for W in "$@"; do python bench.py --steps 2 --warmup 3 --workload ${W%%:*} --samples ${W##*:} --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['config']['workload'][:30], '%.4e'%d['value'], '%.1f ms'%d['ms_per_step'], '%.2f TF %.1f%%'%(r['achieved'],100*r['frac']), 'e2e %.4e'%d['e2e']['value'])
except Exception as e: print('ERR', e)"; done
