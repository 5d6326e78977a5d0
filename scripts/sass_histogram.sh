#!/bin/bash
# Per-kernel histogram of the SASS mnemonics that prove the Blackwell-native path (B200_PROFILING.md): UTC*MMA = tcgen05.mma,
# LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor loads / stores, UBLKCP = bulk copies, LDGSTS = cp.async, HMMA = legacy mma.sync.
cd "$(dirname "$0")/.."
cuobjdump -sass minppo_b200/lib/libminppo_b200.so | awk '
  /Function : /{f=$3}
  /^ *\/\*[0-9a-f]+\*\// {
    op = $2; if (op ~ /^@/) op = $3;
    if (op ~ /^(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|LDGSTS|HMMA|HGMMA|UTCBAR|MUFU\.TANH)/) { split(op, a, "."); key = a[1]; if (op ~ /^MUFU\.TANH/) key = "MUFU.TANH"; c[f" "key]++ } }
  END { for (k in c) print k, c[k] }' | sort | c++filt | sed 's/minppo:://g' | awk '{n=$NF; k=$(NF-1); $NF=""; $(NF-1)=""; printf "%-70s %-10s %s\n", $0, k, n}'
