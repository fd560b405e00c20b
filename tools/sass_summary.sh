#!/bin/bash
# Counts of the SASS mnemonics that prove tcgen05 / TMA / TMEM use, per object and per kernel (B200_PROFILING.md):
#   UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit
# Usage: bash tools/sass_summary.sh > profiles/r2_sass_summary.txt   (after python -m comfy_rvc_b200.build)
cd "$(dirname "$0")/../comfy_rvc_b200/build" || exit 1
echo "# cuobjdump -sass of comfy_rvc_b200/build/*.o (sm_100a), $(date -u +%Y-%m-%d), git $(git rev-parse --short HEAD 2>/dev/null)"
echo "# per object: UTCHMMA UTMALDG UTMASTG LDTM UTCBAR"
for o in *.o; do
  s=$(cuobjdump -sass "$o" 2>/dev/null)
  printf "%-22s %6d %6d %6d %6d %6d\n" "$o" $(echo "$s" | grep -c UTCHMMA) $(echo "$s" | grep -c UTMALDG) $(echo "$s" | grep -c UTMASTG) \
      $(echo "$s" | grep -c LDTM) $(echo "$s" | grep -c UTCBAR)
done
echo
echo "# rmvpe_kernels.o (GRU recurrence over a thread-block cluster): STAS = st.async (DSMEM store + mbarrier complete_tx), SYNCS = mbarrier ops, UCGABAR = barrier.cluster, MAPA"
s=$(cuobjdump -sass rmvpe_kernels.o 2>/dev/null)
printf "%-22s STAS %d  SYNCS %d  UCGABAR %d  MAPA %d\n" rmvpe_kernels.o $(echo "$s" | grep -c "STAS") $(echo "$s" | grep -c "SYNCS") $(echo "$s" | grep -c "UCGABAR") $(echo "$s" | grep -c "MAPA")
echo
echo "# per kernel (objects with tcgen05 code): function, UTCHMMA, UTMALDG, UTMASTG, LDTM"
for o in rbconv_tc.o rbpair_tc.o conv_tc.o attention_tc.o; do
  cuobjdump -sass "$o" 2>/dev/null | awk -v obj="$o" '
    /Function : / { if (name != "") printf "%s %s %d %d %d %d\n", obj, name, a, b, c, d; name=$3; a=b=c=d=0 }
    /UTCHMMA/ {a++} /UTMALDG/ {b++} /UTMASTG/ {c++} /LDTM/ {d++}
    END { if (name != "") printf "%s %s %d %d %d %d\n", obj, name, a, b, c, d }' | while read obj name a b c d; do
      printf "%-14s %-110s %5d %5d %5d %5d\n" "$obj" "$(echo "$name" | c++filt | cut -c1-110)" "$a" "$b" "$c" "$d"
    done
done
