#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU needed) into a small text file for profiles/.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_sweep.txt"""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["ncu summary of %s (ncu --set full --clock-control none; per launch)" % rep, ""]
    for r in rows[2:]:
        lines.append("kernel: " + r[hdr.index("Kernel Name")][:100])
        for k in KEYS:
            if k in hdr:
                lines.append("  %-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        st = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued"):
                try:
                    st.append((float(r[i].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1.0
        lines.append("  warp-state samples: " + ", ".join("%s %.1f%%" % (h, 100 * v / tot) for v, h in sorted(st, reverse=True)[:8]))
        lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
