"""Counts the SASS mnemonics that identify the Blackwell data paths (B200_PROFILING.md, "What proves a Blackwell-native
kernel") per kernel of libradtts_b200.so.  Runs without a GPU:  python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "radtts_b200", "lib", "libradtts_b200.so")
FAMILIES = [
    ("UTC*MMA (tcgen05.mma)", re.compile(r"\bUTC[A-Z]*MMA\b")),
    ("LDTM/STTM (tcgen05.ld/st)", re.compile(r"\b(LDTM|STTM)\b")),
    ("UTMALDG/UTMASTG/UBLKCP (TMA)", re.compile(r"\b(UTMALDG|UTMASTG|UBLKCP)\b")),
    ("SYNCS (mbarrier)", re.compile(r"\bSYNCS\b")),
    ("HMMA (mma.sync)", re.compile(r"\bHMMA\b")),
    ("LDGSTS (cp.async)", re.compile(r"\bLDGSTS\b")),
    ("SHFL", re.compile(r"\bSHFL\b")),
]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None or "/*" not in line:
            continue
        counts[cur]["instructions"] += 1
        for name, rx in FAMILIES:
            if rx.search(line):
                counts[cur][name] += 1
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    print("# %s: %d kernels, arch %s" % (os.path.basename(lib), len(order), ",".join(arch)))
    print("# columns: " + " | ".join(["instr"] + [n for n, _ in FAMILIES]))
    for k in order:
        c = counts[k]
        if not any(c[n] for n, _ in FAMILIES[:5]):
            continue
        name = re.sub(r"\s+", " ", demangle(k))
        name = name[:150] + ("..." if len(name) > 150 else "")
        print("%6d | %s | %s" % (c["instructions"], " | ".join("%4d" % c[n] for n, _ in FAMILIES), name))
    plain = [k for k in order if not any(counts[k][n] for n, _ in FAMILIES[:5])]
    print("# %d further kernels use none of the tensor/TMA/mbarrier paths (SIMT / shuffle kernels)" % len(plain))


if __name__ == "__main__":
    main()
