#!/usr/bin/env python
"""Turn an .ncu-rep (brought back in gpurun_out/) into the compact per-kernel table committed under
profiles/.  Usage: python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/rNN_name.md"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%peak"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%"),
    ("smsp__inst_executed.sum", "warp_instr"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/instr"),
    ("lts__t_sector_hit_rate.pct", "L2_hit_%"),
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    H, U = rows[0], rows[1]
    ki = H.index("Kernel Name")
    cols = [(H.index(m), n) for m, n in WANT if m in H]
    lines = [f"# ncu summary of `{rep}` (`ncu --set full --clock-control none`)", "",
             "| kernel | " + " | ".join(n for _, n in cols) + " |", "|---|" + "---|" * len(cols)]
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")[:40]
        vals = []
        for i, n in cols:
            v = r[i].replace(",", "")
            try:
                f = float(v)
                if n == "time_us" and U[i] == "ns":
                    f /= 1000.0
                if n.endswith("_MB"):
                    u = U[i].lower()
                    f *= {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1e-6)
                vals.append(f"{f:.1f}" if f < 1e6 else f"{f:.3e}")
            except ValueError:
                vals.append(v)
        lines.append(f"| {name} | " + " | ".join(vals) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
