"""Top stall-sample SASS instructions of each kernel in an .ncu-rep (source page).  usage: ncu_hot.py rep [regex] [topN]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else "."; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if not row: continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}; blocks.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"] = row; continue
    cur["rows"].append(row)
for b in blocks:
    if not re.search(pat, b["name"]): continue
    h = {n: i for i, n in enumerate(b["hdr"])}
    si, ai, ni = h["Source"], h["Warp Stall Sampling (All Samples)"], h["Warp Stall Sampling (Not-issued Samples)"]
    tot = sum(int(r[ai]) for r in b["rows"]) or 1
    print(f"\n== {b['name'][:100]}  total samples {tot}, {len(b['rows'])} SASS instrs")
    idx = sorted(range(len(b["rows"])), key=lambda i: -int(b["rows"][i][ai]))[:top]
    for i in sorted(idx):
        r = b["rows"][i]
        print(f"  #{i:5d} {int(r[ai]):7d} ({100*int(r[ai])/tot:5.1f}%)  {r[si].strip()[:90]}")
