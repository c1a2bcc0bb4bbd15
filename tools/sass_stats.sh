#!/bin/bash
# Per-kernel SASS instruction counts and an opcode histogram of one object file (development aid).
# usage: tools/sass_stats.sh build/traverse.o [kernel-substring]
obj=$1; pat=${2:-}
cuobjdump -sass "$obj" | awk -v pat="$pat" '
/Function :/ {name=$3; sub(/.*trace_kernel/,"trace_kernel",name); sub(/.*[0-9]+shade_kernel/,"shade_kernel",name)}
/^ +\/\*[0-9a-f]+\*\/ / { if (pat=="" || index(name,pat)) { cnt[name]++; op=$2; sub(/\..*/,"",op); if (op ~ /^@/) {op=$3; sub(/\..*/,"",op)}; h[name" "op]++ } }
END { for (n in cnt) print cnt[n], n; if (pat!="") for (k in h) print "   ", h[k], k }' | sort -k1,1 -k2,2nr
