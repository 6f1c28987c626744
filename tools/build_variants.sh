#!/bin/bash
# Builds kernel variants of libsfb200 (-D flags) into build/var/lib_<name>.so for tools/exp_match.py (SFB200_LIB selects one).
# usage: tools/build_variants.sh "name:-DFLAG=.. -DFLAG2=.." ...
cd "$(dirname "$0")/.." && mkdir -p build/var
SRC=$(ls strainflair_b200/csrc/*.cu strainflair_b200/csrc/*.cpp)
FL="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC,-pthread -shared -cudart shared -Iinclude -Istrainflair_b200/csrc"
for v in "$@"; do
  n=${v%%:*}; d=${v#*:}
  ( nvcc $FL $d -Xptxas -v -o build/var/lib_$n.so $SRC -lz 2>&1 | grep -A3 "${KERNEL:-k_matchE}" | grep "Used\|spill" | tr '\n' ' '; echo " <- $n ($d)" ) &
done
wait
