#!/bin/bash
# Run every example for two optimizer steps on whatever device is present (CPU works: models fall back to tiny random presets,
# datasets to synthetic ones).  Usage: scripts/smoke_examples.sh [timeout-seconds-per-example]
cd "$(dirname "$0")/.."
export PYTHONPATH="$PWD:${PYTHONPATH:-}" TRLX_B200_OFFLINE=${TRLX_B200_OFFLINE:-1} TRLX_B200_MAX_EVAL_PROMPTS=${TRLX_B200_MAX_EVAL_PROMPTS:-8}
T=${1:-600}
OUT=$(mktemp -d)
HP='{"train.total_steps": 2, "train.batch_size": 4, "train.eval_interval": 2, "train.checkpoint_interval": 1000, "train.tracker": null, "train.checkpoint_dir": "'$OUT'/ckpt"}'
fail=0
for ex in $(find examples -name "*.py" | sort); do
  grep -q "def main(hparams" "$ex" || continue
  case "$ex" in *nemo_vs_ds_chat*|*inference*) continue;; esac
  start=$(date +%s)
  if timeout "$T" python "$ex" "$HP" > "$OUT/log.txt" 2>&1; then status=ok; else status="FAIL($?)"; fail=1; fi
  echo "$status $(( $(date +%s) - start ))s $ex"
  [ "$status" = ok ] || tail -n 3 "$OUT/log.txt" | cut -c1-200
done
exit $fail
