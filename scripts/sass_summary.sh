#!/bin/bash
# SASS evidence per kernel of the shipped library: counts of the Blackwell-native mnemonics
#   UTCHMMA (tcgen05.mma kind::f16), LDTM / STTM (tcgen05.ld / st), UBLKCP (cp.async.bulk, TMA engine), UTCBAR (tcgen05.commit),
#   SYNCS (mbarrier ops), FFMA2 / FADD2 (packed fp32x2), MUFU.* -- and the absence of HMMA (legacy mma.sync) / UTMALDG (tensor-map TMA).
# usage: scripts/sass_summary.sh > profiles/sass_summary.txt
SO=${1:-rlaopt_b200/csrc/librlaopt_b200.so}
echo "# cuobjdump -sass $SO  ($(date -u +%Y-%m-%d), $(sha256sum $SO | cut -c1-16))"
echo "# kernel | mnemonic counts"
cuobjdump -sass $SO 2>/dev/null | awk '
/Function :/ {name=$3}
!/Function/ {
  n=split($0,a," ");
  for(i=1;i<=n;i++) if (a[i] ~ /^(UTCHMMA|UTCQMMA|LDTM|STTM|UBLKCP|UTCBAR|UTMALDG|UTMASTG|HMMA|SYNCS|FFMA2|FADD2|FMUL2|MUFU\.[A-Z0-9]+|LDGSTS|DFMA|BAR|UCGABAR)/) { split(a[i],b,"."); key=(b[1]=="MUFU") ? a[i] : b[1]; gsub(/;$/,"",key); cnt[name"|"key]++; names[name]=1 }
}
END { for (k in cnt) print k, cnt[k] }' | sort | while IFS='|' read -r mangled rest; do echo "$(echo $mangled | c++filt | sed 's/(anonymous namespace):://; s/kmm:://g; s/(.*//') | $rest"; done | awk -F' \\| ' '{k[$1]=k[$1] "  " $2} END {for (n in k) print n " |" k[n]}' | sort
echo
echo "# totals"
cuobjdump -sass $SO 2>/dev/null | grep -o -E "UTCHMMA|LDTM|STTM|UBLKCP[A-Z.]*MULTICAST|UBLKCP|UTCBAR|UTMALDG|HMMA|FFMA2|MUFU\.EX2|MUFU\.SQRT" | sort | uniq -c
