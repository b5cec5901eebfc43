#!/bin/bash
#SBATCH --job-name=trlx_b200
#SBATCH --nodes=1
#SBATCH --ntasks-per-node=1
#SBATCH --gpus-per-node=8
#SBATCH --cpus-per-task=64
#SBATCH --exclusive
#SBATCH --output=%x_%j.out
# SLURM launch (counterpart of the reference's scripts/slurm_train.sh).  One launcher task per node; the launcher starts one
# rank per GPU.  NCCL over NVLink/NVSwitch inside the node; for multi-node jobs the first node is the rendezvous host.
set -e
export MASTER_ADDR=$(scontrol show hostnames "$SLURM_JOB_NODELIST" | head -n 1)
export MASTER_PORT=${MASTER_PORT:-29500}
export NCCL_DEBUG=${NCCL_DEBUG:-WARN}
export NCCL_NVLS_ENABLE=${NCCL_NVLS_ENABLE:-1}
SCRIPT=${SCRIPT:-examples/ppo_sentiments.py}
PRESET=${PRESET:-configs/accelerate/zero2-bf16.yaml}
cd "${SLURM_SUBMIT_DIR:-$(dirname "$0")/..}"
export PYTHONPATH="$PWD:$PYTHONPATH"
srun python -m trlx_b200.launch --config_file "$PRESET" --num_machines "$SLURM_NNODES" --main_process_ip "$MASTER_ADDR" \
     --main_process_port "$MASTER_PORT" --num_processes 8 "$SCRIPT" "${HPARAMS:-{\}}"
