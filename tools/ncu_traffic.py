"""profiles/mlp_traffic.json from an ncu --set full capture of the dominant kernel's launches (tools/profile_step.py):
average dram__bytes_read.sum + dram__bytes_write.sum per launch.  usage: ncu_traffic.py rep.ncu-rep out.json [kernel substring]"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
want = sys.argv[3] if len(sys.argv) > 3 else "mlp_fused"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot, n, per = 0.0, 0, []
for r in data:
    if want not in r[col["Kernel Name"]]:
        continue
    b = sum(float(r[col[k]].replace(",", "")) * scale[units[col[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    per.append({"kernel": r[col["Kernel Name"]][-60:], "dram_bytes": b,
                "time_us": float(r[col["gpu__time_duration.sum"]].replace(",", ""))})
    tot += b
    n += 1
json.dump({"dram_bytes_per_launch": tot / max(n, 1), "launches": n, "source": f"ncu --set full, {rep}", "per_launch": per},
          open(out, "w"), indent=1)
print(tot / max(n, 1), n)
