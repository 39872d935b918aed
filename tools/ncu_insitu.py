"""Per-kernel DRAM traffic of the second ViT forward, measured IN SITU: `ncu --cache-control none --replay-mode application`-free
single-pass metrics, so the L2 keeps what the previous kernel left in it (the default capture flushes the caches before
every kernel and says nothing about producer -> consumer reuse).  usage: ncu_insitu.py launches.csv  -> table on stdout"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
# columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section, Metric Name, Unit, Value
per = {}
order = []
for r in rows:
    kid, name, metric, unit, val = int(r[0]), r[4], r[-3], r[-2], float(r[-1].replace(",", ""))
    if kid not in per:
        per[kid] = {"name": name}
        order.append(kid)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)
    per[kid][metric] = val * scale
def short(n):
    for k in ("gemm_tcgen05_kernel", "vit_attention_tc", "vit_cls_attention", "layernorm384", "normalize_patchify", "rowstats_cast", "write_cls"):
        if k in n:
            if k == "gemm_tcgen05_kernel":
                i = n.index("<")
                return "gemm" + n[i:i + 40].split(">")[0] + ">"
            return k
    return n[:40]
half = len(order) // 2
tot_r = tot_w = tot_t = 0.0
for kid in order[half:]:
    d = per[kid]
    rd, wr, t = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0), d.get("gpu__time_duration.sum", 0.0)
    tot_r += rd; tot_w += wr; tot_t += t
    print(f"{kid:4d} {short(d['name']):44s} read {rd / 1e6:8.1f} MB  write {wr / 1e6:8.1f} MB  {t:8.1f} us")
print(f"second forward: read {tot_r / 1e6:.0f} MB  write {tot_w / 1e6:.0f} MB  total {(tot_r + tot_w) / 1e6:.0f} MB  kernel time {tot_t:.0f} us")
