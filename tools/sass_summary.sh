#!/bin/bash
# Per-kernel counts of the sm_100a instructions that prove the tcgen05 / TMEM / TMA path (B200_PROFILING.md): run anywhere
# (cuobjdump needs no GPU).   tools/sass_summary.sh > profiles/r02_sass_summary.txt
LIB="$(dirname "$0")/../vehicle_counting_b200/libvcb200.so"
echo "# cuobjdump -sass $(basename $LIB) ($(date -u +%Y-%m-%d)): instruction counts per kernel"
echo "# UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG = TMA load (IM2COL = im2col mode),"
echo "# UTMASTG = TMA store, SYNCS = mbarrier, UTMAPF = tensormap prefetch.  UTCHMMA counts include the .2CTA forms, UTMALDG the .IM2COL forms."
cuobjdump -sass "$LIB" | awk '
/Function :/ { fn=$3; next }
{ for (k in pat) if (index($0, pat[k])) cnt[fn, k]++ }
BEGIN { n=split("UTCHMMA.2CTA UTCHMMA LDTM UTCBAR UTMALDG UTMALDG.4D.IM2COL UTMASTG SYNCS UTMAPF", a, " "); for (i=1;i<=n;i++) pat[a[i]]=a[i] }
END {
  for (key in cnt) { split(key, p, SUBSEP); fns[p[1]]=1 }
  for (fn in fns) {
    line = ""
    for (i=1;i<=n;i++) { c = cnt[fn, a[i]]; if (c) line = line sprintf(" %s=%d", a[i], c) }
    if (line != "") print fn ":" line
  }
}' | sed -e 's/_ZN3vcb[0-9]*//' | c++filt 2>/dev/null | sort
