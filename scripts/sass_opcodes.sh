#!/bin/sh
# Per-kernel SASS opcode histogram of the shipped library (cuobjdump needs no GPU): profiles/r2_sass_opcodes.txt
# Evidence lines to look for: LDTM / STTM (tcgen05.ld / st: multipliers in tensor memory), UBLKCP (cp.async.bulk: TMA staging of
# the parameter record), DFMA / SHFL / REDUX of the ADMM iteration.
LIB=${1:-f110-mpc_b200/libf110mpc_b200.so}
echo "cuobjdump -sass $LIB: opcode histogram per kernel (top 24 opcodes; then the tcgen05 / TMA / spill lines)"
cuobjdump -sass "$LIB" 2>/dev/null | awk '
  /Function :/ { name=$3 }
  /^[ \t]+\/\*[0-9a-f]+\*\/[ \t]/ { op=$2; if (op ~ /^@/) op=$3; sub(/;$/, "", op); split(op, a, "."); base=a[1];
    cnt[name "\t" base]++; tot[name]++;
    if (op ~ /^(LDTM|STTM|UBLKCP|UTMALDG|UTCALLOC|LDL|STL|SYNCS|REDUX|CREDUX)/) { split(op, b, " "); ev[name "\t" op]++ } }
  END { for (k in cnt) print k "\t" cnt[k] > "/tmp/sass_cnt.tsv"; for (k in ev) print k "\t" ev[k] > "/tmp/sass_ev.tsv"; for (n in tot) print n "\t" tot[n] > "/tmp/sass_tot.tsv" }'
sort -t"$(printf '\t')" -k2,2nr /tmp/sass_tot.tsv | while IFS="$(printf '\t')" read -r name total; do
  echo
  echo "== $(echo "$name" | c++filt)  ($total instructions)"
  grep -F "$name	" /tmp/sass_cnt.tsv | sort -t"$(printf '\t')" -k3,3nr | head -24 | awk -F'\t' '{printf "%s:%s ", $2, $3} END {print ""}'
  grep -F "$name	" /tmp/sass_ev.tsv | sort -t"$(printf '\t')" -k2,2 | awk -F'\t' '{printf "   %-14s %s\n", $2, $3}'
done
