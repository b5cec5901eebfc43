#!/bin/bash
# Single- or multi-node launch of an example (counterpart of the reference's scripts/accelerate_train_example.sh).
#   scripts/train_example.sh [examples/ppo_sentiments.py] [configs/accelerate/zero2-bf16.yaml] ['{"train.total_steps": 100}']
set -e
SCRIPT=${1:-examples/ppo_sentiments.py}
PRESET=${2:-configs/accelerate/zero2-bf16.yaml}
HPARAMS=${3:-"{}"}
cd "$(dirname "$0")/.."
export PYTHONPATH="$PWD:$PYTHONPATH"
python -c "import __graft_entry__ as g; g.build()"
NGPU=$(python -c "import torch; print(max(torch.cuda.device_count(), 1))")
python -m trlx_b200.launch --config_file "$PRESET" --num_processes "$NGPU" "$SCRIPT" "$HPARAMS"
