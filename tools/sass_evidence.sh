#!/bin/bash
# Counts of the tcgen05 / TMEM / TMA SASS mnemonics per kernel of the built library (the proof the recipe in
# B200_PROFILING.md asks for): UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (TMA engine),
# UTCBAR = tcgen05.commit, SYNCS = mbarrier.  usage: bash tools/sass_evidence.sh > profiles/sass_evidence_r1.txt
lib=${1:-tfnas_b200/lib/libtfnas_b200.so}
echo "# $(date -u +%FT%TZ)  $lib  ($(nvcc --version | tail -2 | head -1))"
echo "# kernel  UTCHMMA  LDTM  UBLKCP  UTCBAR  SYNCS  total_instructions"
cuobjdump -sass $lib | awk '
/Function : /{name=$3}
/^ +\/\*[0-9a-f]+\*\/ +[A-Z@]/{tot[name]++; if ($0 ~ /UTCHMMA/) a[name]++; if ($0 ~ /LDTM/) b[name]++; if ($0 ~ /UBLKCP/) c[name]++; if ($0 ~ /UTCBAR/) d[name]++; if ($0 ~ /SYNCS/) e[name]++}
END{for(k in tot) if (a[k]+b[k]+c[k]+d[k] > 0) printf "%s %d %d %d %d %d %d\n", k, a[k], b[k], c[k], d[k], e[k], tot[k]}' | sort | c++filt | sed 's/(.*) / /' 
