#!/usr/bin/env python
"""ncu raw-page CSV (ncu -i x.ncu-rep --page raw --csv) -> one line per launch with the columns the profiles/ summaries quote."""
import csv
import sys

COLS = [("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__block_size", "blk"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("smsp__inst_executed.sum", "inst")]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"{'kernel':34s}" + "".join(f"{n:>10s}" for c, n in COLS if c in ix))
    print(f"{'':34s}" + "".join(f"{units[ix[c]][:9]:>10s}" for c, n in COLS if c in ix))
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].replace("void ", "").split("(")[0][:33]
        out = f"{name:34s}"
        for c, n in COLS:
            if c not in ix:
                continue
            v = r[ix[c]].replace(",", "")
            try:
                fv = float(v)
                out += f"{fv:10.1f}" if fv < 1e6 else f"{fv:10.3g}"
            except ValueError:
                out += f"{v[:9]:>10s}"
        print(out)
    # stall reasons of every launch: top 3
    st = [(i, h) for i, h in enumerate(hdr) if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    if st:
        print("\ntop warp-stall reasons per launch (warps stalled per issue-active cycle):")
        for r in rows[2:]:
            name = r[ix["Kernel Name"]].replace("void ", "").split("(")[0][:33]
            vals = []
            for i, h in st:
                try:
                    vals.append((float(r[i].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
            vals.sort(reverse=True)
            print(f"{name:34s}" + "  ".join(f"{n} {v:.2f}" for v, n in vals[:4]))


if __name__ == "__main__":
    main()
