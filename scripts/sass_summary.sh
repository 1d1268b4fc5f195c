#!/bin/bash
# SASS evidence of the Blackwell-native instructions per kernel of pyannote_video_b200/libpvb200.so:
#   UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, UTCBAR = tcgen05.commit, HMMA = legacy mma.sync
# usage: bash scripts/sass_summary.sh > profiles/r02_sass_summary.txt
LIB=${1:-pyannote_video_b200/libpvb200.so}
echo "# cuobjdump -sass $LIB | per-kernel instruction counts ($(date -u +%F))"
cuobjdump -sass "$LIB" | awk '
  /Function : / { name=$3; next }
  name != "" {
    if ($0 ~ /UTC[A-Z]*MMA/) mma[name]++
    if ($0 ~ /LDTM/) ldtm[name]++
    if ($0 ~ /STTM/) sttm[name]++
    if ($0 ~ /UTMALDG/) tmal[name]++
    if ($0 ~ /UTMASTG/) tmas[name]++
    if ($0 ~ /UBLKCP/) blk[name]++
    if ($0 ~ /UTCBAR/) bar[name]++
    if ($0 ~ /[^A-Z]HMMA/) hmma[name]++
    if ($0 ~ /SYNCS/) syncs[name]++
    seen[name]=1
  }
  END {
    printf "%-12s %-8s %-6s %-8s %-8s %-7s %-7s %-6s %-6s %s\n", "UTC*MMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "HMMA", "SYNCS", "kernel"
    for (k in seen) printf "%-12d %-8d %-6d %-8d %-8d %-7d %-7d %-6d %-6d %s\n", mma[k], ldtm[k], sttm[k], tmal[k], tmas[k], blk[k], bar[k], hmma[k], syncs[k], k
  }' | (read -r h; echo "$h"; sort -k1,1nr -k10) | c++filt
