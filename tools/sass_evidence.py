"""Per-kernel counts of the SASS mnemonics that identify the Blackwell paths (tcgen05 MMA / commit / TMEM loads, TMA bulk
copies and reductions, mbarriers, fp64 DMMA) in the built library.  Needs only cuobjdump (no GPU).
    python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "egobox_b200", "libegobox_gpu.so")
PAT = re.compile(r"\b(UTCIMMA|UTCHMMA|UTCQMMA|UTCBAR|UTCCP|LDTM|STTM|UBLKCP|UBLKRED|UTMALDG|UTMASTG|UTMAREDG|SYNCS|DMMA|"
                 r"UTCATOMSWS|USETMAXREG)\b")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, counts = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
        elif cur:
            for mn in PAT.findall(line):
                counts[cur][mn] += 1
    names = [k for k in counts if counts[k]]
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.strip().splitlines()
    rows = []
    for raw, name in zip(names, dem):
        name = re.sub(r"^void ", "", name.replace("(anonymous namespace)::", ""))
        depth, cut = 0, len(name)
        for i in range(len(name) - 1, -1, -1):           # drop the argument list
            if name[i] == ")":
                depth += 1
            elif name[i] == "(":
                depth -= 1
                if depth == 0:
                    cut = i
                    break
        rows.append((name[:cut], counts[raw]))
    print("# SASS evidence (cuobjdump -sass egobox_b200/libegobox_gpu.so, sm_100a): instruction counts per kernel\n")
    print("UTCIMMA = tcgen05.mma kind::i8 | UTCBAR = tcgen05.commit | LDTM = tcgen05.ld (tensor memory -> registers) | "
          "UTCATOMSWS = tcgen05.alloc/dealloc")
    print("UBLKCP = cp.async.bulk (1-D TMA) | UBLKRED / UTMAREDG = bulk / tensor-map reductions (measured variants of the C "
          "update, off by default)")
    print("SYNCS = mbarrier operations | USETMAXREG = setmaxnreg | DMMA = fp64 mma.sync (the non-tcgen05 GEMM / solve kernels)\n")
    for name, c in sorted(rows):
        print("%-64s %s" % (name[:64], "  ".join("%s=%d" % kv for kv in sorted(c.items()))))


if __name__ == "__main__":
    sys.exit(main())
