#!/bin/bash
# profiles/rNN_sass_summary.txt: per-kernel SASS mnemonic counts (Blackwell evidence: UBLKCP = cp.async.bulk / TMA,
# SYNCS = mbarrier, LDGSTS = cp.async, FFMA2 / FMUL2 / FADD2 = packed FP32, FMNMX3 = 3-input min/max, DFMA = FP64) from
# `cuobjdump -sass` of the shipped library + registers / spills from `-Xptxas -v` of every translation unit.
LIB=${1:-sac_rcbf_b200/librcbf_b200.so}
echo "# cuobjdump -sass $LIB : instruction and mnemonic counts per kernel"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { if (name != "") flush(); name=$3; n=0; delete c; next }
  /^[[:space:]]+\/\*[0-9a-f]+\*\/[[:space:]]/ {
    n++; line=$0; sub(/^[[:space:]]+\/\*[0-9a-f]+\*\/[[:space:]]+/, "", line); sub(/^@!?U?P[0-9T]+[[:space:]]+/, "", line);
    split(line, f, /[ .;]/); op=f[1]; c[op]++ }
  function flush() { printf "%-110s n=%6d UBLKCP=%d SYNCS=%d LDGSTS=%d FFMA2=%d FMUL2=%d FADD2=%d FMNMX3=%d DFMA=%d FFMA=%d MUFU=%d\n", substr(name,1,110), n, c["UBLKCP"], c["SYNCS"], c["LDGSTS"], c["FFMA2"], c["FMUL2"], c["FADD2"], c["FMNMX3"], c["DFMA"], c["FFMA"], c["MUFU"] }
  END { if (name != "") flush() }' | sed 's/_ZN[0-9]*_GLOBAL__N__[0-9a-f_]*cu_[0-9a-f]*//' | sort
echo
echo "# nvcc -Xptxas -v (registers, spills, shared memory) per entry function"
for f in rcbf_kernels rcbf_safe_unicycle rcbf_safe2_unicycle rcbf_safe_cars rcbf_cars2 rcbf_gp rcbf_general rcbf_replay; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v -c sac_rcbf_b200/csrc/$f.cu -o /tmp/sass_$f.o 2>&1 |
    awk -v tu=$f '/Compiling entry function/ { name=$0; sub(/.*entry function ./, "", name); sub(/. for .*/, "", name) }
         /bytes spill stores/ { if (!seen[name]) { spill=$0; sub(/^[[:space:]]+/, "", spill) } }
         /Used [0-9]+ registers/ { u=$0; sub(/.*Used /, "Used ", u); printf "%-14s %-100s %s | %s\n", tu, substr(name,1,100), u, spill; seen[name]=1 }' &
done
wait
