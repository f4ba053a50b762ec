"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: ncu -i rep --page source --csv --print-source cuda,sass > x.csv; python tools/ncu_lines.py x.csv [top]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
csv.field_size_limit(1 << 30)
cur_file = None; hdr = None; cur_line = None; cur_src = None
agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), ""])   # samples, inst, stalls, src
for r in csv.reader(open(path, errors="replace")):
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] != "":
        cur_line = r[0]; cur_src = r[1].strip(); continue      # the CUDA line row carries aggregated numbers; use SASS rows
    if r[2] in ("...", "-"): continue
    try:
        ns = int(r[hdr.index("# Samples")]); ni = int(r[hdr.index("Instructions Executed")])
    except ValueError:
        continue
    a = agg[(cur_file, cur_line)]
    a[0] += ns; a[1] += ni; a[3] = cur_src
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            try: v = int(r[i])
            except ValueError: v = 0
            if v: a[2][h[6:]] += v
tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
print(f"total samples {tot_s}  total warp-instructions {tot_i}")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ",".join(f"{k}:{v}" for k, v in a[2].most_common(3))
    print(f"{100*a[0]/max(tot_s,1):5.1f}% smp {100*a[1]/max(tot_i,1):5.1f}% ins  {f}:{l:>4s}  [{st}]  {a[3][:90]}")
