#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show a Blackwell-native kernel (B200_PROFILING.md): tcgen05.mma ->
UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG, tcgen05.commit -> UTCBAR, packed fp32 -> FFMA2/FADD2/FMUL2,
3-input minima -> FMNMX3; legacy tensor paths (HMMA, HGMMA) must be absent.

  python tools/sass_counts.py [openvino-sam-6d_b200/libpem_b200.so] > profiles/r2_sass_tc.txt"""
import collections
import re
import subprocess
import sys

WATCH = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP",
         "SYNCS", "FFMA2", "FADD2", "FMUL2", "FMNMX3", "REDUX", "HMMA", "HGMMA", "MUFU.EX2", "ELECT")


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else "openvino-sam-6d_b200/libpem_b200.so"
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for w in WATCH:
                if op.startswith(w):
                    per[cur][w if not op.startswith("UTCHMMA.2CTA") else "UTCHMMA.2CTA"] += 1
                    if ".2CTA" in op and not op.startswith("UTCHMMA"):
                        per[cur][w + ".2CTA"] += 1
    print("# SASS mnemonic counts per kernel of %s (cuobjdump -sass; kernels without any watched mnemonic omitted)" % lib)
    tot = collections.Counter()
    for k, c in per.items():
        if not c:
            continue
        tot.update(c)
        print("%-70s %s" % (k[:70], "  ".join("%s=%d" % kv for kv in sorted(c.items()))))
    print("\nTOTAL " + "  ".join("%s=%d" % kv for kv in sorted(tot.items())))
    assert tot["HMMA"] == 0 and tot["HGMMA"] == 0, "legacy tensor-core path present"


if __name__ == "__main__":
    main()
