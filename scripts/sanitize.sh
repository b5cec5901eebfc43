#!/usr/bin/env bash
# compute-sanitizer targets for the sm_100a kernels: memcheck over a small-shape selection of the
# kernel tests (every kernel family runs at least once), then racecheck + synccheck over the kernels that use plain
# shared-memory hand-offs (decode / RL math / optimizer).  The tcgen05 GEMM family synchronises through mbarriers
# and the async proxy, which racecheck does not model, so it is covered by memcheck + synccheck only.
# Usage: scripts/sanitize.sh [outdir]   (needs a GPU; logs land in <outdir>, default gpurun_out/sanitize)
#        SAN_ONLY="gemm.memcheck plain.racecheck" restricts the passes, SAN_TIMEOUT bounds each one (seconds).
set -u
cd "$(dirname "$0")/.."
OUT=${1:-gpurun_out/sanitize}
mkdir -p "$OUT"
CS=${COMPUTE_SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}
PY="python -m pytest tests/test_kernels_gpu.py -x -q -p no:cacheprovider"
export PYTORCH_NO_CUDA_MEMORY_CACHING=1   # every tensor gets its own cudaMalloc, so OOB accesses are attributable

GEMM_SMALL='test_gemm_matches_fp32 and (32-768-64 or 100-776-3072) or test_gemm_mn_major_operands and 300-200-136 or test_gemm_split_k and 3 or test_gemm_with_folded_norm or test_gemm_cta_pair and 512-512-256 or test_gemm_fp8 and 300- or test_gemm_epilogue or test_gemm_strided_input or test_lmhead_logprob and 77-1000-256 or test_lmhead_greedy'
PLAIN='test_norm or test_embed_rowdot or test_decode_attention or test_logprob_from_logits or test_gae_whiten or test_ppo_loss_and_grads or test_rollout_rewards or test_adamw_flat'
# round-2 kernels: radix-select samplers, ILQL sampler, 8-bit Adam, row-parallel column sums, in-place wgrad accumulation
NEW='test_sample_filtered or test_ilql_sample or test_adam8bit or test_wgrad_accumulates or test_bias_gradient_column_sum'

run() {  # name, tool, selection, seconds
  local name=$1 tool=$2 sel=$3 secs=$4
  if [ -n "${SAN_ONLY:-}" ] && [[ " $SAN_ONLY " != *" $name.$tool "* ]]; then return 0; fi
  echo "== $name ($tool)"
  timeout "$secs" "$CS" --tool "$tool" --error-exitcode 9 --print-limit 20 --launch-timeout 0 \
      --log-file "$OUT/$name.$tool.log" $PY -k "$sel" > "$OUT/$name.$tool.pytest.log" 2>&1
  local rc=$?
  echo "   exit=$rc  $(tail -n 1 "$OUT/$name.$tool.pytest.log")"
  grep -h "ERROR SUMMARY\|RACECHECK SUMMARY" "$OUT/$name.$tool.log" | tail -n 1
  return 0
}

run gemm memcheck "$GEMM_SMALL" ${SAN_TIMEOUT:-420}
run plain memcheck "$PLAIN" ${SAN_TIMEOUT:-420}
run plain racecheck "$PLAIN" ${SAN_TIMEOUT:-420}
run gemm synccheck "$GEMM_SMALL" ${SAN_TIMEOUT:-420}
run new memcheck "$NEW" ${SAN_TIMEOUT:-420}
run new racecheck "$NEW" ${SAN_TIMEOUT:-420}
