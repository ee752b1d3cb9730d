#!/bin/bash
# Opcode evidence from the built library (no GPU needed): TMA bulk copy + mbarrier of the
# scene staging, 128-bit stores, system-scope atomics of the completion counter, no tensor ops.
# usage: tools/sass_evidence.sh > profiles/rNN_sass_evidence.txt
LIB=quadray-engine_b200/lib/libquadray_b200.so
K='_Z16qr_render_kernelILb1ELi896ELi1EEv9qr_launch'
T=$(mktemp -d); (cd $T && cuobjdump -xelf all $OLDPWD/$LIB > /dev/null); C=$(ls $T/*.cubin | head -1)
echo "library: $LIB ($(git rev-parse --short HEAD 2>/dev/null))"; echo "kernel : $K"
nvdisasm -c $C > $T/all.sass
awk -v k=".text.$K" '$0 ~ "^//-+ .text" {p = index($0, k) > 0} p' $T/all.sass > $T/k.sass
echo "instructions: $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' $T/k.sass)"
for op in UBLKCP SYNCS 'STG.E.128' 'LDS.128' 'STS.128' 'LDG.E.64.CONSTANT' 'ATOMG\|RED' 'FFMA' 'FMUL' 'FADD' 'MUFU' 'HMMA\|UTCMMA\|TCGEN\|UTMA' 'LDL\|STL'; do
  printf "%-22s %5d   e.g. %s\n" "$op" "$(grep -c "$op" $T/k.sass)" "$(grep -m1 "$op" $T/k.sass | sed 's/\/\*[0-9a-f]*\*\///g; s/^ *//' | cut -c1-80)"
done
rm -rf $T
