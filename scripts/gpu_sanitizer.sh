# compute-sanitizer over one small rollout per kernel family (memcheck, racecheck, synccheck); logs -> gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
run() {  # tool tag args...
  tool=$1; tag=$2; shift 2
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_run.py "$@" > gpurun_out/sanitizer_${tool}_${tag}.log 2>&1
  echo "== $tool $tag: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_${tag}.log | tail -1) | $(grep sanitize_run gpurun_out/sanitizer_${tool}_${tag}.log | tail -1)"
}
for tool in memcheck racecheck synccheck; do
  run $tool ts_swarm50 swarm50 300 1 f32 tc
  run $tool tc_swap12 swap12 300 2 f32 tc
  run $tool tc_singlequad singlequad 300 2 f32 tc
  run $tool tile_swarm50_f32 swarm50 100 1 f32 tile
  run $tool tile_singlequad_f64 singlequad 300 2 f64 tile
  run $tool vec_softcorridor softcorridor 4 3 f32 vec
done
