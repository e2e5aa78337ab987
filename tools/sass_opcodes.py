#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libdinopose_sm100a.so (cuobjdump -sass): the mnemonics that prove a Blackwell-native
kernel -- UTC*MMA (tcgen05.mma), UTCCP (tcgen05.cp), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tile load / store),
UTCBAR (tcgen05.commit) -- next to HMMA (legacy mma.sync).    python tools/sass_opcodes.py > profiles/r2_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dino_pose_b200", "libdinopose_sm100a.so")
COLS = ["UTCHMMA", "UTCCP", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "HMMA", "MUFU", "total"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(n):
    n = re.sub(r"\(anonymous namespace\)::|dp::|void ", "", n)
    return re.sub(r"\(.*", "", n)[:70]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    fn, hist = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            hist[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and fn:
            op = m.group(1)
            hist[fn]["total"] += 1
            for c in COLS:
                if op.startswith(c):
                    hist[fn][c] += 1
    names = demangle(list(hist))
    rows = [(short(names[f]), h) for f, h in hist.items()]
    tc = [r for r in rows if r[1]["UTCHMMA"]]
    other = [r for r in rows if not r[1]["UTCHMMA"]]
    print(f"# SASS opcode histogram of `{os.path.relpath(SO, ROOT)}` ({len(rows)} kernels; `cuobjdump -sass`, `tools/sass_opcodes.py`)\n")
    print("`UTCHMMA` = tcgen05.mma (kind::f16), `UTCCP` = tcgen05.cp, `UTCBAR` = tcgen05.commit, `LDTM`/`STTM` = tcgen05.ld/st, "
          "`UTMALDG`/`UTMASTG` = TMA tile load / store, `HMMA` = legacy mma.sync.\n")
    tot = collections.Counter()
    for _n, h in rows:
        tot.update(h)
    print("Library totals: " + ", ".join(f"{c} {tot[c]}" for c in COLS[:-1]) + "\n")
    print("## Kernels that issue tcgen05.mma\n")
    print("| kernel | " + " | ".join(COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    for n, h in sorted(tc, key=lambda r: r[0]):
        print(f"| `{n}` | " + " | ".join(str(h[c]) for c in COLS) + " |")
    hm = [r for r in other if r[1]["HMMA"]]
    print(f"\n## Kernels with HMMA and no tcgen05.mma: {len(hm)}\n")
    for n, h in hm:
        print(f"* `{n}`: HMMA {h['HMMA']}")
    print(f"\n## Other kernels (CUDA cores / memory bound): {len(other) - len(hm)}\n")
    print(", ".join(f"`{n}`" for n, _h in sorted(other, key=lambda r: r[0]) if not _h["HMMA"]))


if __name__ == "__main__":
    main()
