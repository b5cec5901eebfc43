#!/bin/bash
# Multi-node launch of the LLaMA model-parallel PPO example (reference: examples/llama_nemo/dist_train.sh, an sbatch script
# that srun's one NeMo process per GPU).  One task per node here; each starts 8 ranks with torch.distributed.run.
#SBATCH --job-name=llama-nemo-ppo
#SBATCH --nodes=4
#SBATCH --ntasks-per-node=1
#SBATCH --gpus-per-node=8
#SBATCH --cpus-per-task=64
#SBATCH --exclusive
#SBATCH --output=%x_%j.out

set -euo pipefail
export MASTER_ADDR=${MASTER_ADDR:-$(scontrol show hostnames "${SLURM_JOB_NODELIST:-localhost}" | head -n 1)}
export MASTER_PORT=${MASTER_PORT:-29500}
export NCCL_DEBUG=${NCCL_DEBUG:-WARN}
NNODES=${SLURM_NNODES:-1}
SCRIPT=${1:-examples/llama_nemo/nemo_llama2_ppo_sentiments.py}

LAUNCH="python -m torch.distributed.run --nnodes=$NNODES --nproc-per-node=8 --rdzv-backend=c10d \
  --rdzv-endpoint=$MASTER_ADDR:$MASTER_PORT $SCRIPT"
if command -v srun >/dev/null 2>&1 && [ "$NNODES" -gt 1 ]; then
  srun --kill-on-bad-exit=1 bash -c "$LAUNCH"
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node="${NPROC:-8}" --master-addr 127.0.0.1 --master-port "$MASTER_PORT" "$SCRIPT"
fi
