"""Dev tool: aggregate an `ncu --page source --csv` dump: stall totals, samples per opcode, hottest instructions."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = collections.Counter(); byop = collections.Counter(); exe = collections.Counter()
top = []
for r in data:
    if len(r) < len(hdr): continue
    src = r[ix['Source']]; toks = src.split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    n = int(r[ix['# Samples']] or 0); ex = int(r[ix['Instructions Executed']] or 0)
    byop[op] += n; exe[op] += ex
    for s in stalls: tot[s] += int(r[ix[s]] or 0)
    top.append((n, r[ix['Address']], src, ex, {s: int(r[ix[s]] or 0) for s in stalls if int(r[ix[s]] or 0) > 0}))
T = sum(byop.values())
print("total samples", T, " total warp-instructions executed", sum(exe.values()))
print("stall totals:", [(k, v) for k, v in tot.most_common(10)])
print("by opcode (samples, executed):")
for op, n in byop.most_common(30): print(f"  {op:44s} {n:7d} {100*n/T:5.1f}%  exec {exe[op]}")
print("top instrs:")
nshow = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for n, a, src, ex, st in sorted(top, reverse=True)[:nshow]:
    print(f"{n:6d} {a[-5:]} ex={ex:8d} {src[:72]:72s} {sorted(st.items(), key=lambda x: -x[1])[:3]}")
