"""One line per captured launch of an .ncu-rep: the metrics the roofline claims rest on.   python profiles/ncu_table.py X.ncu-rep"""
import csv
import subprocess
import sys

COLS = [("Kernel Name", "kernel", 44), ("gpu__time_duration.sum", "us", 9),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %", 13),
        ("dram__bytes_read.sum", "DRAM rd MB", 11), ("dram__bytes_write.sum", "DRAM wr MB", 11),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 7),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM MB", 10), ("lts__t_sector_hit_rate.pct", "L2 hit %", 8),
        ("smsp__issue_active.avg.pct", "issue %", 8), ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %", 6),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 %", 7),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefront %", 16),
        ("smsp__cycles_elapsed.avg.per_second", "SM GHz", 7), ("launch__registers_per_thread", "regs", 5)]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
print("  ".join(t.ljust(w) for _, t, w in COLS))
for r in rows[2:]:
    cells = []
    for name, _, w in COLS:
        if name not in hdr:
            cells.append("-".ljust(w)); continue
        v, u = r[hdr.index(name)], units[hdr.index(name)]
        try:
            f = float(v.replace(",", ""))
            if u in ("byte", "Kbyte", "Gbyte"):
                f = f * {"byte": 1e-6, "Kbyte": 1e-3, "Gbyte": 1e3}[u]
            if u == "ns":
                f /= 1e3
            if u == "ms":
                f *= 1e3
            v = "%.2f" % f
        except ValueError:
            v = v[:w]
        cells.append(v.ljust(w))
    print("  ".join(cells))
