#!/bin/bash
#SBATCH --job-name=trlx_b200-sweep
#SBATCH --nodes=1
#SBATCH --ntasks-per-node=1
#SBATCH --gpus-per-node=8
#SBATCH --exclusive
#SBATCH --output=%x_%j.out
# Cluster sweep (same role and name as the reference's scripts/sweep-cw.sh, which boots a Ray cluster over the allocation).
# Here the sweep driver needs no cluster runtime: it runs on the allocation's first node and starts every trial as its own
# `trlx_b200.launch` job over NUM_GPUS ranks (trials run one after another, each using the whole node).
#   sbatch scripts/sweep-cw.sh [sweep.yml] [script.py] [gpus per trial]
set -euo pipefail
export NCCL_DEBUG=${NCCL_DEBUG:-WARN}
export NCCL_NVLS_ENABLE=${NCCL_NVLS_ENABLE:-1}
cd "${SLURM_SUBMIT_DIR:-$(dirname "$0")/..}"
export PYTHONPATH="$PWD:${PYTHONPATH:-}"
SWEEP=${1:-configs/sweeps/ppo_sweep.yml}
SCRIPT=${2:-examples/ppo_sentiments.py}
NUM_GPUS=${3:-8}
python -c "import __graft_entry__ as g; g.build()"
python -m trlx_b200.sweep -y --config "$SWEEP" --num_gpus "$NUM_GPUS" "$SCRIPT"
# python -m trlx_b200.sweep -y --config configs/sweeps/ilql_sweep.yml --default_config configs/ilql_config.yml --num_gpus "$NUM_GPUS" examples/ilql_sentiments.py
