"""SASS opcode histogram of the conv engine (and, with --all, every pv2 kernel) in the built library: the evidence that the convs
run on tcgen05 (UTCHMMA / UTCBAR / UTMALDG / UTMAPF / LDTM / SYNCS) and not on mma.sync (HMMA).  CPU only.
    python profiles/sass_histogram.py > profiles/r2_sass_conv.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pranet-v2_b200", "libpranetv2_b200.so")
WANT = re.compile(r"conv_fwd2_kernel|conv_fwd_kernel|conv_wgrad_kernel") if "--all" not in sys.argv else re.compile(r"pv2")
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UTMAPF", "UTMASTG", "LDTM", "STTM", "SYNCS", "HMMA", "IMMA", "RED", "ATOM", "ATOMG", "LDG", "STG", "LDS", "STS",
       "BAR", "ELECT", "FENCE", "ACQBULK", "MUFU", "FFMA", "DADD", "DFMA")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    cur, hist = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            cur = name if WANT.search(name) else None
            if cur:
                hist[cur] = collections.Counter()
            continue
        if cur:
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m:
                hist[cur][m.group(1)] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a); opcode counts per kernel (base mnemonic, modifiers dropped)")
    for name, c in hist.items():
        short = re.sub(r"\(anonymous namespace\)::", "", name)
        short = re.sub(r"\(.*", "", short)
        total = sum(c.values())
        print(f"\n== {short}: {total} instructions")
        print("   key: " + ", ".join(f"{k} {c[k]}" for k in KEY if c[k]))
        print("   top: " + ", ".join(f"{k} {v}" for k, v in c.most_common(14)))
        if "conv" in short:
            assert c["HMMA"] == 0, "mma.sync in a conv kernel"


if __name__ == "__main__":
    main()
