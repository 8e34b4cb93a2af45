"""Summarise an .ncu-rep (ncu --set full) into a small CSV: one row per captured kernel, the metrics the roofline talk needs.

usage: python tools/ncu_summary.py gpurun_out/<tag>_prof.ncu-rep profiles/<name>.csv
"""
import csv, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "duration_us", 1e-3),
    ("dram__bytes_read.sum", "dram_read_B", None),
    ("dram__bytes_write.sum", "dram_write_B", None),
    ("lts__t_sectors.sum", "l2_sectors", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct", 1),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "mem_throughput_pct", 1),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct", 1),
    ("sm__inst_executed_pipe_tc.sum", "tc_pipe_inst", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct", 1),
    ("sm__inst_executed_pipe_uniform.sum", "uniform_pipe_inst", 1),
    ("smsp__inst_executed.sum", "warp_inst", 1),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__grid_size", "grid", 1),
    ("launch__block_size", "block", 1),
    ("launch__shared_mem_per_block_dynamic", "smem_dyn", 1),
    ("launch__waves_per_multiprocessor", "waves", 1),
    ("sm__cycles_active.max", "sm_cycles_active_max", 1),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units = rows[0], rows[1]
    kn = h.index("Kernel Name")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [k[1] for k in KEYS])
        for r in rows[2:]:
            vals = []
            for name, _, scale in KEYS:
                if name in h:
                    v = r[h.index(name)].replace(",", "")
                    u = units[h.index(name)]
                    try:
                        x = float(v)
                        if scale == 1e-3: x *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(u, 1.0)
                        if "byte" in u.lower():
                            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                            x *= mult
                        vals.append("%.6g" % x)
                    except ValueError:
                        vals.append(v)
                else:
                    vals.append("")
            w.writerow([r[kn][:80]] + vals)
    print(open(out).read())


if __name__ == "__main__":
    main()
