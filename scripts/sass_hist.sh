#!/bin/bash
# usage: scripts/sass_hist.sh <function-substring> [lib] [top]  -- opcode histogram of one kernel's SASS
LIB=${2:-sac_rcbf_b200/librcbf_b200.so}
cuobjdump -sass "$LIB" | sed -n "/Function : .*$1/,/\.\.\.\.\.\.\.\.\.\./p" | grep -E "^\s+/\*[0-9a-f]{4,}\*/\s" > /tmp/one.sass
echo "instructions: $(grep -c . /tmp/one.sass)"
sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/^@!?U?P[0-9T]+\s+//' /tmp/one.sass | awk '{print $1}' | sed 's/\..*//; s/;//' | sort | uniq -c | sort -rn | head -${3:-45}
