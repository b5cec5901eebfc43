#!/bin/bash
# Fails when the tree that travels to a GPU box (everything except .git/, gpurun_out/ and .gpurunignore entries) is larger
# than a limit (default 100 MB).  Round 1 lost all of its driver-side measurements to a 498 MB example checkpoint in the tree.
cd "$(dirname "$0")/.."
LIMIT_MB=${1:-100}
excl=(--exclude=.git --exclude=gpurun_out)
while IFS= read -r line; do
  line=${line%/}; [ -z "$line" ] && continue; case "$line" in \#*) continue;; esac
  excl+=(--exclude="$line")
done < .gpurunignore
kb=$(du -sk "${excl[@]}" . | cut -f1)
mb=$((kb / 1024))
echo "snapshot size: ${mb} MB (limit ${LIMIT_MB} MB)"
if [ "$mb" -gt "$LIMIT_MB" ]; then
  echo "largest entries:"; du -sk "${excl[@]}" ./* | sort -rn | head -5
  exit 1
fi
