"""Markdown table from an .ncu-rep (ncu -i ... --page raw --csv): one row per captured launch.

    python tools/ncu_summary.py gpurun_out/r01b_attn.ncu-rep [more.ncu-rep ...]
"""
import csv
import io
import subprocess
import sys

COLS = [
    ("time µs", "gpu__time_duration.sum", lambda v, u: f"{float(v) * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}.get(u, 1):.1f}"),
    ("DRAM rd MB", "dram__bytes_read.sum", lambda v, u: f"{float(v) * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}.get(u, 1):.1f}"),
    ("DRAM wr MB", "dram__bytes_write.sum", lambda v, u: f"{float(v) * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}.get(u, 1):.1f}"),
    ("DRAM % peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", lambda v, u: f"{float(v):.1f}"),
    ("SM % peak", "sm__throughput.avg.pct_of_peak_sustained_elapsed", lambda v, u: f"{float(v):.1f}"),
    ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", lambda v, u: f"{float(v):.1f}"),
    ("issue slots %", "sm__inst_issued.avg.pct_of_peak_sustained_active", lambda v, u: f"{float(v):.1f}"),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", lambda v, u: f"{float(v):.1f}"),
    ("warp inst M", "smsp__inst_executed.sum", lambda v, u: f"{float(v) / 1e6:.1f}"),
    ("regs", "launch__registers_per_thread", lambda v, u: f"{int(float(v))}"),
    ("grid x block", None, None),
]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    for row in r[2:]:
        yield hdr, units, row


def main():
    print("| kernel | " + " | ".join(c[0] for c in COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    for rep in sys.argv[1:]:
        for hdr, units, row in rows_of(rep):
            def get(name):
                if name in hdr:
                    i = hdr.index(name)
                    return row[i], units[i]
                return None, None
            name = get("Kernel Name")[0] or "?"
            name = name.split("(")[0].replace("void ", "").replace("kon::", "").replace("unnamed>::", "").replace("<unnamed>::", "")
            cells = []
            for title, metric, fmt in COLS:
                if metric is None:
                    gx, bx = get("Grid Size")[0], get("Block Size")[0]
                    cells.append(f"{gx} x {bx}")
                    continue
                v, u = get(metric)
                try:
                    cells.append(fmt(v.replace(",", ""), u) if v not in (None, "") else "-")
                except Exception:
                    cells.append(str(v))
            print(f"| `{name[:60]}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
