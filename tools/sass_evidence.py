#!/usr/bin/env python3
"""Per-kernel SASS mnemonic counts of the built objects (cuobjdump -sass): which kernels use the TMA bulk-copy unit
(UBLKCP = cp.async.bulk), mbarriers (SYNCS), the integer dot-product pipe (IDP = dp4a / dp2a), shared / global atomics and
reductions, warp votes / shuffles / match, FP64 -- and that none uses a tensor-core instruction (HMMA / IMMA / UTC*MMA: nothing
on this path is a dense contraction).  Static counts, not executed instructions.

    python tools/sass_evidence.py > profiles/r02_sass_evidence.txt
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ("UBLKCP", "SYNCS", "IDP", "ATOMS", "ATOMG", "RED", "REDG", "VOTE", "SHFL", "MATCH", "REDUX", "LDS", "STS", "LDG", "STG", "LDL", "STL",
         "DFMA", "DMUL", "DADD", "MUFU", "HMMA", "IMMA", "UTCHMMA", "UTCIMMA", "UTCQMMA", "BAR")


def main():
    rows = []
    for obj in sorted(glob.glob(os.path.join(ROOT, "leaffliction_b200", "build", "*.o"))):
        out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        name, cnt, total = None, None, 0
        for line in out.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                if name:
                    rows.append((os.path.basename(obj), name, total, cnt))
                sym = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                name = re.sub(r"\(.*", "", sym.replace("(anonymous namespace)::", "")).replace("void ", "").strip()
                cnt, total = collections.Counter(), 0
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m and name:
                total += 1
                op = m.group(1)
                for w in WATCH:
                    if op == w or op.startswith(w + "."):
                        cnt[w] += 1
        if name:
            rows.append((os.path.basename(obj), name, total, cnt))
    print("# static SASS instruction counts per kernel (sm_100a), built objects of leaffliction_b200/csrc")
    print("# kernel | instructions | " + " ".join(WATCH))
    tensor = 0
    for obj, name, total, cnt in rows:
        tensor += cnt["HMMA"] + cnt["IMMA"] + cnt["UTCHMMA"] + cnt["UTCIMMA"] + cnt["UTCQMMA"]
        print(f"{obj[:-2]}:{name} | {total} | " + " ".join(f"{w}={cnt[w]}" for w in WATCH if cnt[w]))
    print(f"# tensor-core instructions in the whole library: {tensor}")


if __name__ == "__main__":
    main()
