"""Summarise an .ncu-rep (read here on the CPU box with `ncu -i`) into the few per-launch metrics the roofline
uses.  usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.summary.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__cycles_elapsed.max", "cycles"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: {len(data)} launches (ncu --set full --clock-control none; cold-cache, serialised replays)")
    for r in data:
        name = r[col["Kernel Name"]]
        print(f"\n{name[:110]}")
        for key, short in KEYS:
            if key in col:
                print(f"  {short:10s} {r[col[key]]:>16s} {units[col[key]]}")


if __name__ == "__main__":
    main(sys.argv[1])
