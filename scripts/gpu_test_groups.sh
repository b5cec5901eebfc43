#!/bin/bash
# run each GPU kernel test function in its own process so a device-side hang is attributed to one group
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for t in test_gemm_fp8_with_row_and_channel_scales test_gemm_cta_pair test_gemm_with_folded_norm_and_row_moments test_gemm_split_k test_gemm_mn_major_operands test_gemm_matches_fp32 test_gemm_persistent_many_tiles test_lmhead_dlogits test_gemm_epilogue test_gemm_strided_input test_lmhead_logprob test_decode_attention test_logprob_from_logits test_gae_whiten test_ppo_loss_and_grads test_kl_rewards test_adamw_flat test_linear_autograd test_fused_logprob_autograd test_lmhead_greedy_and_sampling test_embed_rowdot test_norm; do
  echo "=== $t"
  timeout 100 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "$t" --timeout=45 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -25
  echo "rc=$?"
done
