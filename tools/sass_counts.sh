#!/bin/bash
# per-kernel counts of the Blackwell tensor / TMA instructions in libatvs.so (sm_100a SASS):
#   UTCHMMA = tcgen05.mma.kind::f16, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UTMALDG = TMA tensor load,
#   UBLKCP = cp.async.bulk, LDGSTS = cp.async, SYNCS = mbarrier ops.   bash tools/sass_counts.sh > profiles/rNN_sass_tcgen05_counts.txt
cd "$(dirname "$0")/.."
SO=a-tvsnet_b200/libatvs.so
echo "# $(cuobjdump -lelf $SO | head -3 | tr '\n' ' ')"
echo "# UTCHMMA LDTM STTM UTCBAR UTMALDG UBLKCP LDGSTS SYNCS  kernel"
cuobjdump -sass $SO | awk '
BEGIN { n = split("UTCHMMA LDTM STTM UTCBAR UTMALDG UBLKCP LDGSTS SYNCS", a, " ") }
function flush() { if (name != "" && c["UTCHMMA"] + c["LDTM"] + c["STTM"] + c["UTMALDG"] + c["UBLKCP"] > 0) { line = ""; for (i = 1; i <= n; ++i) line = line sprintf("%6d ", c[a[i]] + 0); print line, name } }
/Function :/ { flush(); name = $3; delete c; next }
{ for (i = 1; i <= n; ++i) if (index($0, a[i])) c[a[i]]++ }
END { flush() }' | while read c1 c2 c3 c4 c5 c6 c7 c8 n; do printf "%6d %6d %6d %6d %6d %6d %6d %6d  %s\n" $c1 $c2 $c3 $c4 $c5 $c6 $c7 $c8 "$(echo $n | c++filt | sed 's/(anonymous namespace):://g; s/(.*//; s/^void //')"; done | sort -k9
