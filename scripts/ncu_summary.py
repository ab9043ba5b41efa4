"""Summarise an .ncu-rep (read here on the CPU box) into a small text file for profiles/.
    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_name.txt "title"
"""
import csv, io, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "inst_executed", "sass__inst_executed_global_loads",
        "sass__inst_executed_global_stores", "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    lines = [f"# {title}", f"# source: {rep} (ncu --set full --clock-control none; cold-cache, serialised replays)", ""]
    for r in rows[2:]:
        lines.append("== " + r[idx["Kernel Name"]])
        for w in WANT:
            if w in idx:
                lines.append(f"  {w:72s} {r[idx[w]]:>20s} {units[idx[w]]}")
        byte_scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
        rd = float(r[idx["dram__bytes_read.sum"]]) * byte_scale.get(units[idx["dram__bytes_read.sum"]], 1)
        wr = float(r[idx["dram__bytes_write.sum"]]) * byte_scale.get(units[idx["dram__bytes_write.sum"]], 1)   # own unit per column
        t = float(r[idx["gpu__time_duration.sum"]]) * {"ms": 1e-3, "us": 1e-6, "s": 1, "ns": 1e-9}.get(units[idx["gpu__time_duration.sum"]], 1e-3)
        lines.append(f"  -> DRAM traffic {(rd + wr) / 1e9:.3f} GB in {t * 1e3:.3f} ms = {(rd + wr) / t / 1e9:.0f} GB/s (under the profiler)")
        top = sorted(((float(r[idx[k]]), k) for k in stall), reverse=True)[:6]
        lines.append("  warps stalled per issue: " + ", ".join(f"{k.split('stalled_')[1].replace('_per_issue_active.ratio', '')} {v:.2f}" for v, k in top))
        lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
