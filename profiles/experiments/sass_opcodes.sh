#!/bin/bash
# Opcode histogram per kernel of the built library (cuobjdump -sass): the Blackwell-specific
# mnemonics (UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor load, UBLKCP =
# cp.async.bulk, FFMA2 = packed fp32 FMA, LDGSTS = cp.async, CREDUX = redux.sync) per kernel.
LIB=${1:-gnn_ecommerce_b200/liblgc_b200.so}
cuobjdump -sass "$LIB" | awk '
/Function :/ { fn=$3; sub(/^_ZN3lgc[0-9]+_GLOBAL__N__[0-9a-f_]+cu_[0-9a-f]+/, "", fn); next }
/^[ \t]+\/\*[0-9a-f]{4}\*\// { op=$2; if (op ~ /^@/) op=$3; sub(/;$/, "", op); split(op, a, "."); base=a[1];
  if (op ~ /^(UTCHMMA|UTCMMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UBLKPF|FFMA2|LDGSTS|CREDUX|UTCBAR|SYNCS|LDG\.E\.128|STG\.E\.128|LDS\.128|STS\.128|SHFL|REDUX)/) { k=base; if (op ~ /^LDG\.E\.128/) k="LDG.128"; if (op ~ /^STG\.E\.128/) k="STG.128"; if (op ~ /^LDS\.128/) k="LDS.128"; if (op ~ /^STS\.128/) k="STS.128"; cnt[fn" "k]++ }
  tot[fn]++ }
END { for (f in tot) { line=f" total="tot[f]; for (key in cnt) { split(key, p, " "); if (p[1]==f) line=line" "p[2]"="cnt[key] } print line } }' | sort
