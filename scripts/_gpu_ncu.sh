# ncu --set full captures (one GPU): GEMM family after the warp-uniform issue loops, fused LM head, optimizer kernels
set -u
timeout 400 bash scripts/ncu_capture.sh gemm "gemm_" 9 2>&1 | tail -3
mv gpurun_out/ncu_gemm_summary.csv gpurun_out/ncu_gemm_v6_summary.csv
timeout 300 bash scripts/ncu_capture.sh lmhead "gemm_tn|lmhead_reduce" 4 2>&1 | tail -3
timeout 300 bash scripts/ncu_capture.sh optim "adam" 4 2>&1 | tail -3
rm -f gpurun_out/*_source.csv gpurun_out/ncu_*.ncu-rep   # keep the summaries + raw pages only (size)
ls -la gpurun_out | grep ncu_ | head
cat gpurun_out/ncu_optim_summary.csv | cut -c1-600
