#!/usr/bin/env python
"""Regenerate profiles/sass_evidence.txt: SASS mnemonic counts of the shipped library and, per kernel,
the tensor-core / TMA / system-scope-atomic instructions it contains (cuobjdump -sass, no GPU needed).

    python tools/sass_evidence.py > profiles/sass_evidence.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "distkeras_b200", "lib", "libdistkeras_b200.so")
INTERESTING = re.compile(r"^(UTC|UTMA|LDTM|STTM|SYNCS|ACQBULK|UCGABAR|REDG|ATOMG|LDGSTS|ARRIVES|UBLKCP|FENCE\.VIEW\.ASYNC)")
PER_KERNEL = [("UTCHMMA", r"^UTCHMMA"), ("UTCHMMA.2CTA", r"^UTCHMMA.*2CTA"), ("LDTM", r"^LDTM"), ("UTMALDG", r"^UTMALDG"),
              ("UTMASTG/REDG", r"^UTMA(STG|REDG)"), ("SYS-atomics", r"^(REDG|ATOMG).*\.SYS"), ("FENCE.ASYNC", r"^FENCE\.VIEW\.ASYNC"), ("LDGSTS", r"^LDGSTS"),
              ("LDGSTSBAR", r"^ARRIVES\.LDGSTSBAR")]


def main() -> None:
    sass = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    total = collections.Counter()
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            if INTERESTING.match(op):
                total[op] += 1
            per[cur][op] += 1
    print(f"# SASS mnemonics in distkeras_b200/lib/libdistkeras_b200.so (cuobjdump -sass, sm_100a); regenerate with tools/sass_evidence.py")
    print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG/UTMASTG/UTMAREDG = cp.async.bulk.tensor load/store/reduce,")
    print("# UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc, FENCE.VIEW.ASYNC = fence.proxy.async, REDG/ATOMG ... .SYS = system-scope PS atomics")
    for op, n in total.most_common():
        print(f"{n:7d} {op}")
    print("\n# kernels containing tensor-core / TMA / system-scope instructions")
    for name, c in per.items():
        cols = []
        for label, pat in PER_KERNEL:
            n = sum(v for k, v in c.items() if re.match(pat, k))
            if n:
                cols.append(f"{label}={n}")
        if cols:
            short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
            print(f"{short[:90]:90s} {' '.join(cols)}")


if __name__ == "__main__":
    sys.exit(main())
