#!/bin/bash
# Opcode evidence for DESIGN.md section 5: per translation unit, the tcgen05 / TMEM / TMA / legacy-MMA SASS mnemonics in the built objects.
#   bash scripts/sass_histogram.sh > profiles/r02_sass_opcodes.txt
echo "# cuobjdump -sass of boxdreamer_b200/_build/*.o (sm_100a), counts of the mnemonics that identify the data path"
echo "# UTCHMMA = tcgen05.mma (kind::f16), .2CTA = cta_group::2; LDTM / STTM = tcgen05.ld / .st; UTMALDG / UTMASTG / UTMAREDG = cp.async.bulk.tensor"
echo "# load / store / reduce; UTCBAR = tcgen05.commit; HMMA = legacy mma.sync (only the DINOv2 prefix-row side kernel); MUFU.EX2 = exp2"
for o in boxdreamer_b200/_build/*.o; do
  echo
  echo "## $(basename $o)"
  cuobjdump -sass $o | grep -oE "\b(UTCHMMA(\.2CTA)?|UTCQMMA|UTCBAR(\.2CTA)?(\.MULTICAST)?|LDTM(\.x[0-9]+)?|STTM(\.x[0-9]+)?|UTMALDG\.[0-9]D(\.2CTA)?|UTMASTG\.[0-9]D|UTMAREDG\.[0-9]D\.ADD|UTMAPF|UTMACCTL\.PF|UBLKCP|HMMA\.[0-9]+\.F32(\.BF16)?|MUFU\.EX2|MUFU\.RSQ|MUFU\.RCP|SYNCS\.[A-Z_.]+|UTCATOMSWS[A-Z_.0-9]*|FFMA2|DFMA|DMUL|DADD)\b" | sort | uniq -c | sort -k1,1nr | awk '{printf "  %-40s %s\n", $2, $1}'
done
echo
echo "## per kernel (functions with tcgen05 MMAs / TMA / TMEM traffic)"
for o in boxdreamer_b200/_build/gemm_tc2.o boxdreamer_b200/_build/attn_tc2.o; do
  cuobjdump -sass $o | awk '/Function :/ {fn=$NF} /UTCHMMA/ {m[fn]++} /UTMALDG|UTMASTG|UTMAREDG/ {t[fn]++} /LDTM|STTM/ {q[fn]++} END {for (f in m) printf "%s UTCHMMA=%d TMA=%d LDTM+STTM=%d\n", f, m[f], t[f], q[f]}' | c++filt | sed 's/^/  /' | sort
done
