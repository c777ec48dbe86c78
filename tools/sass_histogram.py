#!/usr/bin/env python
"""Opcode histogram of selected kernels from the shipped library's SASS (cuobjdump -sass), for profiles/.
    python tools/sass_histogram.py ithaca-sem_b200/libnekmf_b200.so 'hex_helm_kron_kernelILi5ELi2ELb0ELi0' ... > profiles/rNN_sass_histogram.txt
Each argument after the library is a substring of a mangled kernel name."""
import collections
import re
import subprocess
import sys


def main():
    lib, pats = sys.argv[1], sys.argv[2:]
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, hist = None, {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1) if any(p in m.group(1) for p in pats) else None
            if cur:
                hist[cur] = collections.Counter()
            continue
        if cur:
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                hist[cur][m.group(1)] += 1
    for fn, h in hist.items():
        demangled = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip() or fn
        tot = sum(h.values())
        print("==== %s\n     %s\n     %d instructions" % (demangled[:150], fn[:120], tot))
        groups = collections.Counter()
        for op, n in h.items():
            groups[op.split(".")[0]] += n
        for op, n in groups.most_common(24):
            print("  %-10s %6d  %5.1f %%" % (op, n, 100.0 * n / tot))
        for key in ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "DFMA", "DMMA", "LDGSTS", "LDS", "STS", "UTCMMA", "HMMA"):
            print("  [%s: %d]" % (key, sum(n for op, n in h.items() if op.startswith(key))), end="")
        print()


if __name__ == "__main__":
    main()
