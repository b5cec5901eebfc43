#!/bin/bash
# Run the reference example set on the CURRENT tree with the jsonl tracker (learning-curve regression runs).
# Counterpart of the reference's scripts/benchmark.sh (clone + venv + W&B); here: local tree, offline, jsonl logs.
#   scripts/benchmark.sh [--logs DIR] [--only_hash] [--only_tiny]
set -e
logs=benchmark_logs/current
only_hash=false
only_tiny=false
while [[ "$#" -gt 0 ]]; do
    case $1 in
        --logs) logs="$2"; shift ;;
        --only_hash) only_hash=true ;;
        --only_tiny) only_tiny=true ;;
        --origin|--branch) shift ;;   # accepted for CLI compatibility: the tree in front of us is what gets measured
        --public) ;;
        *) echo "Unknown parameter passed: $1"; exit 1 ;;
    esac
    shift
done
here="$(cd "$(dirname "$0")/.." && pwd)"
cd "$here"
if [ "$only_hash" = true ]; then
    python -c "from trlx_b200.reference import content_hash; print(content_hash('.'))"
    git log --format=%h/%s/%as -n1 2>/dev/null || true
    exit 0
fi
mkdir -p "$logs"
run() {  # run <name> <script> [launcher args...]
    name=$1; script=$2; shift 2
    args='{"train.tracker": "jsonl", "train.logging_dir": "'$logs/$name'", "train.checkpoint_dir": "'$logs/$name/ckpts'"}'
    PYTHONPATH="$here" "$@" "$script" "$args" > "$logs/$name.log" 2>&1 || echo "$name failed (see $logs/$name.log)"
}
[ -f python_build_done ] || python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1 || true
run ilql_randomwalks examples/randomwalks/ilql_randomwalks.py python
run ppo_randomwalks examples/randomwalks/ppo_randomwalks.py python
if [ "$only_tiny" = true ]; then exit 0; fi
ngpu=$(python -c "import torch; print(torch.cuda.device_count())")
if [ "$ngpu" -lt 1 ]; then echo "no GPU: skipping the sentiment experiments"; exit 0; fi
i=0
for ex in ppo_sentiments sft_sentiments ilql_sentiments ppo_sentiments_t5; do
    CUDA_VISIBLE_DEVICES=$((i % ngpu)) run $ex examples/$ex.py python &
    i=$((i + 1))
    if [ $((i % ngpu)) -eq 0 ]; then wait; fi
done
wait
