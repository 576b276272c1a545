#!/bin/bash
# usage: scripts/sass_hist.sh object.o kernel_substring  -> opcode histogram of that kernel's SASS
cuobjdump -sass "$1" 2>/dev/null | awk '/Function : /{name=$3} {print name "\t" $0}' | grep "$2" | awk -F'\t' '{print $2}' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's#/\*[0-9a-f]+\*/##g' | awk '{op=$1; if (op ~ /^@/) op=$2; split(op,a,"."); c[a[1]]++; n++} END{for(k in c) print c[k], k; print n, "TOTAL"}' | sort -rn | head -${3:-30}
