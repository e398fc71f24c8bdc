#!/bin/bash
# compute-sanitizer sweep (GPU box): memcheck + racecheck + synccheck on small plans of every kernel family.
mkdir -p gpurun_out
OUT=gpurun_out/sanitizer.txt
: > $OUT
run() {  # tool, label, command...
  tool=$1; label=$2; shift 2
  echo "== $tool : $label" >> $OUT
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 "$@" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard|RESULT|passed|failed" | head -12 >> $OUT
}
# racecheck does not model mbarrier arrive / try_wait between warps: the kernels with the W exchange report their
# publish / consume pairs as hazards unless built with -DNDNLM_DEBUG_CTA_SYNC=1 (NDNLM_EXTRA_NVCC_FLAGS); by default
# racecheck only runs on the kernels without that exchange.  RACE_ALL=1 runs it everywhere (slow: ~10 min).
for tool in racecheck; do
  run $tool "nlm tiled 2-D f=1"                       python tools/dev_parity.py --case 4
  run $tool "box mean (reference_compiled fast path)" python tools/dev_parity.py --case 13
  if [ -n "$RACE_ALL" ]; then
    run $tool "nlm tiled f=1 cfg3-like (TMA, 2 passes)" python tools/dev_parity.py --case 2
    run $tool "nlm tiled float64 (T = double)"          python tools/dev_parity.py --case 10
  fi
done
for tool in memcheck synccheck; do
  run $tool "nlm tiled f=1 cfg3-like (TMA, 2 passes)" python tools/dev_parity.py --case 2
  run $tool "nlm tiled 2-D f=1"                       python tools/dev_parity.py --case 4
  run $tool "nlm tiled n_eff"                         python tools/dev_parity.py --case 11
  run $tool "nlm tiled f=2"                           python tools/dev_parity.py --case 12
  run $tool "nlm tiled float64 (T = double)"          python tools/dev_parity.py --case 10
  run $tool "box mean (reference_compiled fast path)" python tools/dev_parity.py --case 13
done
run memcheck "sibling filters (all kernels)" python -m pytest tests/test_sibling_filters.py -q -m gpu -x
run memcheck "nlm gpu parity subset" python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "matches_oracle_float32 or sharded or slab"
cat $OUT
