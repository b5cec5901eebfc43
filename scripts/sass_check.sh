#!/bin/bash
# Proof that the built extension contains Blackwell tensor-core / TMA / cluster instructions, per kernel family
# (B200_PROFILING.md: UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA load / store, LDTM = tcgen05.ld, UCGABAR = cluster barrier).
# Usage: scripts/sass_check.sh [path/to/_C.so]      (runs without a GPU)
cd "$(dirname "$0")/.."
SO=${1:-trlx_b200/_C.so}
[ -f "$SO" ] || { echo "build first: python -c 'import __graft_entry__ as g; g.build()'"; exit 1; }
cuobjdump -sass "$SO" 2>/dev/null | awk '
  /Function : / { name=$3; sub(/^_ZN4b200[0-9]*/, "", name); sub(/I[LE].*/, "", name); sub(/E[vP].*/, "", name) }
  { for (i = 1; i <= NF; i++) if ($i ~ /^(UTCHMMA|UTCQMMA|UTMALDG|UTMASTG|LDTM|UTCBAR|UCGABAR_ARV|SYNCS)/) { k = $i; sub(/\..*/, "", k); c[name " " k]++ } }
  END { for (x in c) print c[x], x }' | sort -k2,2 -k3,3 | awk '{printf "%-34s %-12s %s\n", $2, $3, $1}'
echo "--- registers / stack / shared per kernel"
cuobjdump -res-usage "$SO" 2>/dev/null | grep -A1 "Function" | grep -oE "Function [^:]*|REG:[0-9]+|STACK:[0-9]+|SHARED:[0-9]+" | paste - - - - | sed 's/Function _ZN4b200[0-9]*//' | cut -c1-110 | sort | head -60
