"""Static SASS instruction mix of every kernel in the objects of astr_b200/csrc (cuobjdump -sass): the mnemonics that
prove which data path a kernel uses (UTMALDG = TMA tensor load, UBLKCP = bulk copy, SYNCS = mbarrier, LDGSTS =
cp.async, DFMA/DADD/DMUL = fp64 pipe, STL/LDL = local memory).
usage: python tools/sass_summary.py astr_b200/csrc/*.o > profiles/<name>.txt"""
import re
import subprocess
import sys
from collections import Counter

KEYS = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "DFMA", "DADD", "DMUL", "MUFU", "BAR",
        "LDL", "STL", "ATOM", "RED", "BRA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main(paths):
    rows = []
    for path in paths:
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        cur, cnt, tot = None, None, 0
        for line in sass.split("\n"):
            m = re.search(r"Function : (\S+)", line)
            if m:
                if cur:
                    rows.append((path, cur, tot, cnt))
                cur, cnt, tot = m.group(1), Counter(), 0
                continue
            m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(@!?\S+\s+)?([A-Z][A-Z0-9_]*)", line)
            if cur and m:
                tot += 1
                op = m.group(2)
                for k in KEYS:
                    if op == k or (k in ("LDG", "STG", "LDS", "STS", "LDL", "STL", "BAR", "ATOM", "RED") and op.startswith(k)):
                        cnt[k] += 1
                        break
        if cur:
            rows.append((path, cur, tot, cnt))
    names = demangle([r[1] for r in rows])
    print("# static SASS instruction mix per kernel (cuobjdump -sass, sm_100a)")
    print(f"{'kernel':78s} {'instr':>6s} " + " ".join(f"{k:>7s}" for k in KEYS))
    for path, fn, tot, cnt in rows:
        n = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", names.get(fn, fn))
        n = re.sub(r"\(.*", "", n).replace("void ", "")
        print(f"{n[:78]:78s} {tot:6d} " + " ".join(f"{cnt.get(k, 0):7d}" for k in KEYS))


if __name__ == "__main__":
    main(sys.argv[1:])
