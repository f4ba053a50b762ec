"""Summarise ncu outputs under gpurun_out/ into profiles/ (tracked)."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
launch_csv = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
rep = os.path.join(ROOT, "gpurun_out", sys.argv[2] if len(sys.argv) > 2 else f"prof_{tag}_inv.ncu-rep")
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list ----
rows = []
with open(launch_csv) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((r["Kernel Name"], ns))
agg = {}
for k, ns in rows:
    k = k.split("(")[0]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += ns
tot = sum(a[1] for a in agg.values())
with open(os.path.join(out_dir, f"{tag}_launch_list_summary.md"), "w") as f:
    f.write(f"# ncu launch list ({tag}) — `python bench.py --draws 16 --points 131072 --steps 2 --warmup 1 --no-cpu-baseline`\n\n")
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (cold-cache, serialised: compare SHARES).\n\n")
    f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {ns/1e6:.3f} | {100*ns/tot:.1f} % |\n")
    f.write(f"\nTotal {tot/1e6:.2f} ms over {len(rows)} launches.  The pack kernels run once per draw set (outside the timed region);\n"
            "inside a timed step the launches are: 1 x flow_tc_inv_kernel + 1 x lse_finish_kernel.\n")

# ---- full capture of the dominant kernel ----
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
lines = [l for l in raw.splitlines() if l and not l.startswith("==")]
rr = list(csv.reader(lines))
hdr, units, vals = rr[0], rr[1], rr[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum"]
def tobytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
with open(os.path.join(out_dir, f"{tag}_flow_tc_inv_kernel_ncu.md"), "w") as f:
    f.write(f"# ncu --set full, flow_tc_inv_kernel ({tag}) — 16 draws x 131072 points, config-3 flow\n\n| metric | value | unit |\n|---|---:|---|\n")
    for w in want:
        if w in m:
            f.write(f"| {w} | {m[w][0]} | {m[w][1]} |\n")
traffic = None
if "dram__bytes_read.sum" in m:
    traffic = tobytes(*m["dram__bytes_read.sum"]) + tobytes(*m["dram__bytes_write.sum"])
    evals = 16 * 131072
    with open(os.path.join(out_dir, f"{tag}_flow_tc_inv_kernel_ncu.md"), "a") as f:
        f.write(f"\nDRAM traffic of this launch: {traffic/1e6:.1f} MB = {traffic/evals:.1f} B per eval "
                f"(algorithmic minimum 20 B/eval for x, ctx and the [N] output; packed weights of 16 draws = 125 MB are read once).\n")
print("traffic bytes", traffic)
