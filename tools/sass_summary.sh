#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove the Blackwell path (B200_PROFILING.md): tcgen05.mma -> UTC*MMA,
# tcgen05.ld/st -> LDTM/STTM, cp.async.bulk -> UBLKCP, tcgen05.commit -> UTCBAR, mbarrier -> SYNCS.*
#   tools/sass_summary.sh > profiles/r2_sass_summary.txt
so=pytorch_bayesiancnn_b200/libbbb_b200.so
echo "# cuobjdump -sass $so (built $(date -u +%F) from $(git rev-parse --short HEAD)); instruction counts per kernel"
cuobjdump -sass $so | awk '
/Function :/ { fn=$3; next }
{ for (i=1;i<=NF;i++) { t=$i; if (t ~ /^(UTCHMMA|UTCQMMA|UTCBAR|LDTM|STTM|UBLKCP|UTMALDG|UTMASTG|SYNCS|HMMA|LDGSTS|UTCATOMSWS|STG|LDG|ATOMG|REDG)/) { split(t,a,"."); c[fn" "a[1]]++ } } }
END { for (k in c) print k, c[k] }' | sort | c++filt | awk '{ k=$0; sub(/ [A-Z]+ [0-9]+$/,"",k); } { print }' | sed 's/void bbb:://' | cut -c1-160
