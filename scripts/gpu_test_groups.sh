#!/bin/bash
# Run every GPU kernel test FUNCTION in its own process, so a device-side hang or a sticky CUDA error is attributed to one group
# instead of poisoning the rest of the file.  (The normal way to run the suite is `python -m pytest tests -m gpu`.)
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
groups=$(grep -oE "^def (test_[A-Za-z0-9_]+)" tests/test_kernels_gpu.py | awk '{print $2}')
for t in $groups; do
  echo "=== $t"
  timeout 120 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "$t" --timeout=60 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -3
  echo "rc=$?"
done
