#!/bin/bash
# compute-sanitizer sweep over the kernels added in round 2 (tcgen05 convolutions + stem, conv-epilogue statistics ->
# BatchNorm apply, attention forward / backward, fused LayerNorm, packed FedPM vote, device-stream Gaussian noise,
# ghost-norm DP step).  Small shapes only; CUDA-graph tests are excluded (the sanitizer's legacy-stream activity
# invalidates stream capture -- an API artefact, see profiles/sanitizer_r1.md).
#   gpurun --timeout 1800 -- 'bash tools/sanitize_r2.sh > gpurun_out/sanitize_r2.log 2>&1'
set -u
cd "$(dirname "$0")/.."
out=${OUT_DIR:-gpurun_out/sanitize_r2}; mkdir -p $out
TESTS="tests/test_gpu_conv.py tests/test_gpu_attention.py tests/test_gpu_layer_norm.py tests/test_gpu_dp.py tests/test_gpu_flat_ops.py"
FILTER='(ragged_batch or l4_1x1_s2 or l2_1x1_s2 or stem_conv or epilogue_statistics or (forward_and_backward_match and (16 or 77)) or (without_dropout and 256) or ghost_clipping or (packed_vote and not 1048576)) and not graph and not resnet and not replay'
PER_TOOL_TIMEOUT=${PER_TOOL_TIMEOUT:-420}
status=0
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  echo "=== compute-sanitizer --tool $tool ==="
  FL4H_NO_AUTOBUILD=1 timeout $PER_TOOL_TIMEOUT compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 20 \
    python -m pytest $TESTS -m gpu -q -k "$FILTER" -p no:cacheprovider -p no:faulthandler > "$out/sanitize_$tool.full.log" 2>&1
  rc=$?
  grep -E "^=========|passed|failed|deselected" "$out/sanitize_$tool.full.log" | grep -v "^=========     " | tail -15 | cut -c1-220
  echo "--- $tool exit code: $rc"
  [ "$rc" -ne 0 ] && status=$rc
done
exit $status
