"""Per-kernel SASS mnemonic counts of libkon_b200.so (cuobjdump -sass): which kernels are tcgen05 / TMEM / bulk-copy
code, which are legacy warp-MMA, which are plain SIMT.  Runs without a GPU.

    python tools/sass_summary.py > profiles/r02_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ml_function_b200", "libkon_b200.so")
WATCH = ["UTCHMMA", "UTCBAR", "STTM", "LDTM", "UTCATOMSWS", "UBLKCP", "SYNCS", "HMMA", "LDSM", "FFMA2", "HMUL2",
         "MUFU", "LDG", "STG", "LDS", "STS", "ATOM", "RED", "BAR", "SHFL", "ERRBAR", "MEMBAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kern = None
    counts = collections.OrderedDict()
    total = collections.Counter()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = m.group(1)
            counts[kern] = collections.Counter()
            continue
        if kern is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            total[kern] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    counts[kern][w] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("# SASS summary of `ml_function_b200/libkon_b200.so` (sm_100a), `cuobjdump -sass`\n")
    print("`UTCHMMA` = tcgen05.mma, `STTM`/`LDTM` = tcgen05.st/ld (tensor memory), `UTCBAR` = tcgen05.commit, `UBLKCP` = "
          "cp.async.bulk (TMA engine, 1-D), `SYNCS` = mbarrier, `HMMA`/`LDSM` = legacy warp MMA / ldmatrix, `FFMA2`/`HMUL2` = "
          "packed fp32 / bf16 math.  Columns with a zero count are left blank.\n")
    cols = [w for w in WATCH if any(c[w] for c in counts.values())]
    print("| kernel | instr | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    seen = {}
    for (k, c), name in zip(counts.items(), demangle):
        base = re.sub(r"<.*", "", re.sub(r"\(anonymous namespace\)::", "", re.sub(r"^void ", "", name)))
        sig = (base, tuple(c[w] for w in cols))
        if sig in seen:                 # template instances with identical instruction mix: listed once
            seen[sig] += 1
            continue
        seen[sig] = 1
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"\(.*", "", name)
        name = name.replace("kon::", "")
        if name.startswith(("cub::", "thrust::", "void cub", "_ZN3cub")) or "cub::" in name:
            name = "cub::" + re.sub(r".*?(Device\w+Kernel).*", r"\1", name)
        print(f"| `{name[:70]}` | {total[k]} | " + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |")


if __name__ == "__main__":
    sys.exit(main())
