#!/bin/bash
# Hyper-parameter sweep over the node's GPUs (counterpart of the reference's scripts/sweep-cw.sh, which starts a Ray cluster).
#   scripts/sweep.sh configs/sweeps/ppo_sweep.yml examples/ppo_sentiments.py [gpus per trial]
set -e
cd "$(dirname "$0")/.."
export PYTHONPATH="$PWD:$PYTHONPATH"
python -m trlx_b200.sweep --config "${1:-configs/sweeps/ppo_sweep.yml}" --num_gpus "${3:-1}" -y "${2:-examples/ppo_sentiments.py}"
