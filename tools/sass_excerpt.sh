#!/bin/bash
# SASS evidence for profiles/: which memory / atomic instructions the hot kernels really use (cuobjdump on the built library, no GPU needed).
# Usage: bash tools/sass_excerpt.sh > profiles/r2_sass_excerpt.txt
LIB=mlir-hashjoin_b200/lib/libhashjoin_b200.so
echo "# cuobjdump -sass $LIB (sm_100a): per kernel, how many of each memory / atomic / barrier instruction"
cuobjdump -sass "$LIB" 2>/dev/null | c++filt | awk '
  /Function :/ { name=$0; sub(/.*Function : /,"",name); sub(/\(.*/,"",name) }
  /ATOMG|ATOMS|REDG|LDG\.E|STG\.E|LDS|STS|UBLKCP|CCTL|BAR\.SYNC|SHFL|VOTE|MATCH|LDGSTS/ {
    op=$2; if (op ~ /^@/) op=$3; sub(/;.*/,"",op); n=split(op,p,"."); key=p[1]; for(i=2;i<=n&&i<=4;i++) key=key"."p[i]; c[name" | "key]++ }
  END { for (k in c) print c[k], k }' | sort -t'|' -k1,1 -k2,2 | awk '{cnt=$1; $1=""; print cnt"\t"$0}' | grep -E "k_build_hash<long, true>|k_count<int, true, 0u>|k_count<long, true, 0u>|k_write_range<int, true>|k_rp_scatter<long, 1, false, 1024, 8>|k_rp_hist<long, 1>|k_rj_join<long, false>|k_rj_join<long, true>|k_rj_emit|k_count_dense_tma|k_group_count<long, true>" | sort -k2
