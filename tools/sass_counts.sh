#!/bin/bash
# Per-kernel SASS mnemonic counts that prove which hardware path a kernel uses (read here, no GPU needed):
#   UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP / UTMALDG = TMA bulk / tensor copies, UTCBAR = tcgen05.commit,
#   DMMA = FP64 tensor pipe, SYNCS = mbarrier.   Usage: tools/sass_counts.sh > profiles/sass_counts_rNN.txt
for o in build/obj/capi_tc.*.o build/obj/capi_tc_i8.*.o build/obj/capi_envdmma.*.o build/obj/capi.*.o; do
  echo "== $o"
  cuobjdump -sass "$o" | python3 -c '
import sys, re, collections
cnt = collections.Counter(); tot = collections.Counter(); name = None
keep = re.compile(r"^(UTC\w*MMA|LDTM|STTM|UBLKCP|UTMALDG|UTCBAR|DMMA|SYNCS|UTCCP|LDGSTS|UTCATOMSWS)$")
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m: name = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        tot[name] += 1
        if keep.match(m.group(1)): cnt[(name, m.group(1))] += 1
for n in sorted(tot):
    ops = ", ".join("%s x%d" % (k[1], v) for k, v in sorted(cnt.items()) if k[0] == n)
    if ops: print("  %-70s %6d instructions: %s" % (n[:70], tot[n], ops))
'
done
