#!/bin/bash
# Blackwell-specific SASS in libdpq.so, per kernel of this repository: tensor-core MMA (UTCHMMA),
# TMEM loads (LDTM), TMA bulk copies (UBLKCP), mbarrier synchronisation (SYNCS.*), vector shared-memory
# loads.  Usage: tools/sass_excerpt.sh > profiles/r2_sass_excerpt.txt
HERE="$(cd "$(dirname "$0")/.." && pwd)"
SO="$HERE/deltapq_b200/libdpq.so"
echo "# cuobjdump -sass $(basename "$SO") ($(stat -c %s "$SO") bytes), sm_100a; counts of selected mnemonics per kernel"
cuobjdump -sass "$SO" 2>/dev/null | awk '
  /Function : /{fn=$3; next}
  /^[ \t]*\/\*[0-9a-f]+\*\//{
     m=$2; if (m ~ /^@/) m=$3;
     if (fn ~ /^_ZN3dpq/ && m ~ /^(UTCHMMA|UTCQMMA|UTCBAR|LDTM|STTM|UBLKCP|UTMALDG|SYNCS|LDS\.128|LDS\.64|ATOMS|CCTL|UTCCP|FMNMX3|LDGSTS|REDUX|DFMA)/) { c[fn" "m]++ }
  }
  END{for(k in c) print c[k], k}' | sort -k2,2 -k3,3 | while read n fn m; do printf "%6d  %-28s %s\n" "$n" "$m" "$(echo "$fn" | c++filt | cut -c1-90)"; done
echo
echo "# first tensor-core / TMEM / TMA instructions of gt_tc_filter2_kernel, the table load of scan8_kernel<8>, the MMA group and TMEM loads of encode_tc_kernel<16>"
cuobjdump -sass "$SO" 2>/dev/null | awk '/Function : /{fn=$3} fn ~ /gt_tc_filter2_kernel/ && /UTCHMMA|LDTM|UBLKCP|SYNCS/{print "gt_tc_filter2: " $0}' | grep -v "^\s*$" | cut -c1-140 | grep "UTCHMMA\|LDTM\|UBLKCP" | head -14
cuobjdump -sass "$SO" 2>/dev/null | awk '/Function : /{fn=$3} fn ~ /scan8_kernelILi8E/ && /UBLKCP|SYNCS|LDS.128/{print "scan8<8>: " $0}' | cut -c1-140 | head -14
cuobjdump -sass "$SO" 2>/dev/null | awk '/Function : /{fn=$3} fn ~ /encode_tc_kernelILi16E/ && /UTCHMMA|LDTM|UTCBAR/{print "encode_tc<16>: " $0}' | cut -c1-140 | head -10
