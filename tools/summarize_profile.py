"""Text summary of an ncu report (`--set full`) for profiles/: selected raw metrics of every
profiled launch + per-source-line aggregation of the hottest kernel.

    python tools/summarize_profile.py gpurun_out/prof.ncu-rep contact_map_b200/libcmb200.so KERNEL_MANGLED_SUBSTR > profiles/x.txt
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    rep, so, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    print("# %s" % rep)
    for r in rows[2:]:
        print("\n## launch: %s  grid %s block %s" % (r[hdr.index("Kernel Name")][:90],
                                                   r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        for i, h in enumerate(hdr):
            if h in WANT or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                print("%-78s %-14s %s" % (h, rows[1][i], r[i]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    tmp = "/tmp/_ncu_src.csv"
    open(tmp, "w").write(src)
    print("\n## per-source-line aggregation (tools/ncu_by_line.py)")
    out = subprocess.run([sys.executable, __file__.replace("summarize_profile.py", "ncu_by_line.py"), tmp, so,
                          kernel, "30"], capture_output=True, text=True)
    print(out.stdout)


if __name__ == "__main__":
    main()
