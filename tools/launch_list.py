"""Condense an `ncu --metrics gpu__time_duration.sum --csv` log of bench.py into the launch list of ONE eager step
(kernel, grid, block, µs), plus per-kernel totals.

    python tools/launch_list.py gpurun_out/r02_xdeepfm_launches.csv profiles/r02_xdeepfm_step_launches.csv
"""
import collections
import csv
import re
import sys


def short(n):
    m = re.search(r"(kon::(?:\(anonymous namespace\)::|<unnamed>::)?\w+(?:<[^>(]*>)?)", n)
    if m:
        return m.group(1).replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("kon::", "")
    m = re.search(r"(nvjet_\w+|cub::\w+::\w+|Device\w+Kernel|multi_tensor_apply_kernel|gemmk1_kernel|splitKreduce_kernel)", n)
    if m:
        return m.group(1)
    m = re.search(r"at::native::(?:\(anonymous namespace\)::)?(\w+)", n)
    if m:
        f = re.search(r"(\w+(?:Functor|_kernel_cuda|Op|_impl))", n[n.find(m.group(1)) + len(m.group(1)):])
        return "at::" + m.group(1) + (":" + f.group(1) if f else "")
    return re.sub(r"\(.*", "", n)[:60]


def main(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    seq = []
    for row in r:
        if len(row) <= vi:
            continue
        v = float(row[vi].replace(",", ""))
        v = v / 1e3 if row[ui] == "ns" else (v * 1e3 if row[ui] == "ms" else v)
        seq.append((short(row[ki]), row[gi], row[bi], v))
    # one step = between the last two launches of the first embedding gather of a step
    marks = [i for i, s in enumerate(seq) if s[0].startswith("embed_fwd_vec_kernel")]
    # the isolated back-to-back timing pass launches the gather many times in a row: keep marks followed by other kernels
    marks = [i for i in marks if i + 1 < len(seq) and not seq[i + 1][0].startswith("embed_fwd_vec_kernel")]
    step = seq[marks[-2]:marks[-1]] if len(marks) >= 2 else seq
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(["#", "kernel", "grid", "block", "us"])
        for i, (n, g, b, v) in enumerate(step):
            w.writerow([i, n, g, b, f"{v:.1f}"])
        agg = collections.OrderedDict()
        for n, g, b, v in step:
            a = agg.setdefault(n, [0, 0.0])
            a[0] += 1
            a[1] += v
        tot = sum(v for *_, v in step)
        w.writerow([])
        w.writerow(["# totals of the step", f"{len(step)} launches", "", "", f"{tot:.1f}"])
        for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow(["", n, f"x{c}", f"{100 * v / tot:.1f}%", f"{v:.1f}"])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
