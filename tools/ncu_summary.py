"""Not a test: one line per kernel launch of an `ncu --set full` report (read with
`ncu -i REPORT --page raw --csv`): launch shape, duration, DRAM bytes, throughput percentages and the
three largest warp-stall reasons.  Writes the tables committed as profiles/r02_ncu_attention_kernels.txt.

    python tools/ncu_summary.py gpurun_out/r2_attn_tc2.ncu-rep [more reports...] > profiles/...txt"""
import csv
import io
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__block_size", "blk"),
        ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dsmem"),
        ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue%"),
        ("lts__t_sector_hit_rate.pct", "L2hit%")]
LABELS = [lbl for _, lbl in COLS]
LABELS.insert(7, "DRAM GB/s")
STALL = "smsp__average_warps_issue_stalled_"


def main(paths):
    for path in paths:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                             text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        ci = {n: i for i, n in enumerate(hdr)}
        stall_cols = [n for n in hdr if n.startswith(STALL) and n.endswith("_per_issue_active.ratio")]
        print(f"# {path}")
        print("kernel".ljust(34) + "".join(lbl.rjust(9) if len(lbl) < 9 else " " + lbl for lbl in LABELS) + "   top stalls (warps per issue)")
        for r in rows[2:]:
            name = r[ci["Kernel Name"]].replace("<unnamed>::", "").split("(")[0][:33]
            vals = []
            for n, lbl in COLS:
                if n not in ci or r[ci[n]] == "":
                    vals.append("-")
                    continue
                v = float(r[ci[n]].replace(",", ""))
                u = units[ci[n]]
                if lbl in ("rdMB", "wrMB"):
                    v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                if lbl == "us":
                    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0,
                          "msecond": 1e3}.get(u, 1.0)
                if lbl == "dsmem":
                    v *= {"byte": 1e-3, "Kbyte": 1.0, "Mbyte": 1e3}.get(u, 1.0)
                vals.append(f"{v:.1f}" if v < 1e5 else f"{v:.0f}")
            try:   # measured DRAM traffic / duration
                gbs = (float(vals[5]) + float(vals[6])) / float(vals[0]) * 1e3
                vals.insert(7, f"{gbs:.0f}")
            except ValueError:
                vals.insert(7, "-")
            st = sorted(((float(r[ci[n]] or 0), n[len(STALL):-len("_per_issue_active.ratio")])
                         for n in stall_cols), reverse=True)[:3]
            print(name.ljust(34) + "".join(v.rjust(9) for v in vals) + "   " +
                  ", ".join(f"{n} {v:.2f}" for v, n in st))
        print()


if __name__ == "__main__":
    main(sys.argv[1:])
