#!/bin/bash
# SASS of the hot kernels of libsfb200.so (cuobjdump, no GPU needed): full listing of each kernel + a count of the
# mnemonics that matter for the design (shared-memory atomics, L2 REDs, 256-bit loads, prefetches, warp votes).
# usage: tools/sass_summary.sh <out prefix>      e.g. profiles/r02q_sass
cd "$(dirname "$0")/.."
LIB=strainflair_b200/libsfb200.so
OUT=${1:-profiles/sass}
cuobjdump -sass $LIB > /tmp/sfb_all.sass
: > ${OUT}_summary.txt
for k in k_match_steps k_classifyILb1 k_tile_scatter1 k_tile_share k_pass2_resolveILb1 k_pass2_scatter k_path_reduce; do
  awk -v k="$k" '/Function :/ {on = index($0, k) > 0} on {print}' /tmp/sfb_all.sass > /tmp/sfb_k.sass
  n=$(grep -c "^\s*/\*[0-9a-f]\{4\}\*/" /tmp/sfb_k.sass)
  echo "== $k: $n instructions" >> ${OUT}_summary.txt
  grep -o "^\s*/\*[0-9a-f]\{4\}\*/\s*\(@!\?U\?P[0-9T] \)\?[A-Z0-9_.]*" /tmp/sfb_k.sass | awk '{print $NF}' | \
    grep -E "^(ATOMS|ATOM|RED|REDUX|LDG|LDS|STS|STG|LD\.|ST\.|CCTL|MATCH|VOTE|SHFL|POPC|BAR|DFMA|DMUL|DADD|MUFU|F2I|I2F|LDL|STL|UBLKCP|UTMALDG|LDGSTS)" | \
    sed 's/\.STRONG\.[A-Z]*//; s/\.CONSTANT//' | sort | uniq -c | sort -rn | awk '{printf "   %6d %s\n", $1, $2}' >> ${OUT}_summary.txt
  if [ "$k" = k_tile_share ] || [ "$k" = k_match_steps ]; then cp /tmp/sfb_k.sass ${OUT}_$k.sass; fi
done
echo "written ${OUT}_summary.txt"
