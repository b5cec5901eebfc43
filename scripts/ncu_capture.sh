#!/bin/bash
# One `ncu --set full` capture of a kernel family on ONE GPU (never under a multi-rank launch: ncu replays each kernel ~40x),
# then the raw / source pages as CSV next to the report.  Numbers printed by a process under ncu are never benchmark values.
# Usage: scripts/ncu_capture.sh <gemm|csk|dlogits|lmhead|all> [kernel-regex] [count]
set -euo pipefail
cd "$(dirname "$0")/.."
WHAT=${1:-gemm}; REGEX=${2:-gemm_}; COUNT=${3:-3}
OUT=gpurun_out/ncu_${WHAT}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:${REGEX}" -c "${COUNT}" -f -o "${OUT}" python scripts/ncu_target.py "${WHAT}"
ncu -i "${OUT}.ncu-rep" --page raw --csv > "${OUT}_raw.csv"
ncu -i "${OUT}.ncu-rep" --page source --csv > "${OUT}_source.csv" || true
python - "$OUT" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1] + "_raw.csv")))
hdr = rows[0]
keep = [i for i, h in enumerate(hdr) if h in ("Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__cluster_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active")
        or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"))]
with open(sys.argv[1] + "_summary.csv", "w", newline="") as fh:
    w = csv.writer(fh)
    for r in rows:
        w.writerow([r[i] for i in keep])
print("summary:", sys.argv[1] + "_summary.csv  (copy what should be judged into profiles/)")
PY
