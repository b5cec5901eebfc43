#!/bin/bash
# Per-node half of a multi-node launch (same role and name as the reference's scripts/accelerate_train_example.sh): run it on
# every node with HOSTNAMES ("node0 node1 ..."), MASTER_ADDR, MASTER_PORT and COUNT_NODE exported by the submitting script.
# The node's rank is its position in HOSTNAMES; one process per local GPU is started by trlx_b200.launch.
#   scripts/accelerate_train_example.sh [preset.yaml] [script.py] ['{"train.total_steps": 100}']
set -euo pipefail
PRESET=${1:-configs/accelerate/zero2-bf16.yaml}
SCRIPT=${2:-examples/ilql_sentiments.py}
HPARAMS=${3:-"{}"}
cd "$(dirname "$0")/.."
export PYTHONPATH="$PWD:${PYTHONPATH:-}"
H=$(hostname)
RANK=0; i=0
for h in ${HOSTNAMES:-$H}; do
  if [ "$h" = "$H" ]; then RANK=$i; fi
  i=$((i + 1))
done
NGPU=$(python -c "import torch; print(max(torch.cuda.device_count(), 1))")
exec python -m trlx_b200.launch --config_file "$PRESET" --num_processes "$NGPU" --num_machines "${COUNT_NODE:-1}" \
     --machine_rank "$RANK" --main_process_ip "${MASTER_ADDR:-127.0.0.1}" --main_process_port "${MASTER_PORT:-29500}" \
     "$SCRIPT" "$HPARAMS"
