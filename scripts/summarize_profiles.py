"""Turn the raw artefacts of scripts/gpu_profile_r1.sh (gpurun_out/) into the tracked summaries under profiles/.
Usage: python scripts/summarize_profiles.py r1"""
import collections, csv, os, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = open(f"profiles/{tag}_summary.md", "w")
def P(*a):
    print(*a, file=out)

P(f"# {tag}: ncu / trace summaries (B200, bench workload: one joint-training step at batch 256 -- see scripts/gpu_profile_r2.sh)\n")
# ---- launch list ----------------------------------------------------------------------------------------
src = f"gpurun_out/launches_{tag}.csv"
if os.path.exists(src):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID")
    d = collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault(r[ii], {"k": r[ki]})[r[mi]] = float(r[vi].replace(",", ""))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for v in d.values():
        agg[v["k"][:90]][0] += 1; agg[v["k"][:90]][1] += v.get("gpu__time_duration.sum", 0) / 1e3
    tot = sum(a[1] for a in agg.values())
    P("## Launch list of ONE step (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised)\n")
    P(f"total {tot:.1f} us over {len(d)} launches\n")
    P("| us | launches | share | kernel |\n|---:|---:|---:|---|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:24]:
        P(f"| {a[1]:.1f} | {a[0]} | {100*a[1]/tot:.1f}% | `{k}` |")
    os.system(f"cp {src} profiles/launches_{tag}.csv")
# ---- ncu --set full ---------------------------------------------------------------------------------------
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max"]
for rep in sorted(f for f in os.listdir("gpurun_out") if f.endswith(f"_{tag}.ncu-rep")):
    txt = subprocess.run(["ncu", "-i", f"gpurun_out/{rep}", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    P(f"\n## `ncu --set full --clock-control none` : {rep}\n")
    if rep.startswith("exec_"):
        tot = [float(r[hdr.index("dram__bytes_read.sum")]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_read.sum")]]
               + float(r[hdr.index("dram__bytes_write.sum")]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_write.sum")]]
               for r in rows[2:]]
        import json
        json.dump({"exec_kernel_bytes_per_launch": sum(tot) / len(tot), "launches": len(tot), "per_launch": tot,
                   "source": f"ncu --set full, {rep}"}, open(f"profiles/{tag}_traffic.json", "w"))
    for r in rows[2:]:
        P(f"### `{r[hdr.index('Kernel Name')][:100]}` (launch id {r[0]})\n")
        P("| metric | value | unit |\n|---|---:|---|")
        for k in keys:
            if k in hdr:
                P(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        P("")
for f in (f"trace_{tag}.txt", f"profile_step_{tag}.txt", f"joint_timeline_{tag}.txt", f"pg_step_trace_{tag}.txt"):
    if os.path.exists(f"gpurun_out/{f}"):
        P(f"\n## {f}\n\n```")
        P(open(f"gpurun_out/{f}").read()[-9000:])
        P("```")
out.close()
print(open(f"profiles/{tag}_summary.md").read()[:3000])
