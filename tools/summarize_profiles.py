"""Turns the raw ncu outputs a `tools/r2_profiles.sh <tag>` run left in gpurun_out/ into the small committed
summaries under profiles/: launch list with per-kernel shares, key metrics of the --set full captures, and
profiles/k1_traffic.json (DRAM bytes per K1 launch, read by bench.py for roofline.traffic).
Usage: python tools/summarize_profiles.py <tag> [round-name]"""
import csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r1"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sm__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.max",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "derived__memory_l1_wavefronts_shared_excessive",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed_pipe_tensor.sum", "smsp__cycles_active.avg"]

# ---- launch list
src = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(src):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    agg = {}
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v_us
    ours = {k: v for k, v in agg.items() if "nplda" in k or "::" in k and "at::" not in k}
    tot = sum(v[1] for v in ours.values())
    with open(os.path.join(P, f"{rnd}_launches.csv"), "w") as f:
        f.write(f"# ncu launch list, {rnd} (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)\n")
        f.write("# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --skip-e2e\n")
        f.write("# kernels of libnplda.so only (torch's synthetic-data generator kernels omitted); share = of libnplda time\n")
        f.write("share_pct,launches,avg_us,kernel\n")
        for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{100 * v[1] / tot:.2f},{v[0]},{v[1] / v[0]:.1f},\"{k[:110]}\"\n")
    print("wrote launches", len(ours))

# ---- launch list of the training steps (tools/quick_train.py 131072: NeuralPlda step, DPlda forward, DPlda step)
src = os.path.join(G, f"{tag}_train_launches.csv")
if os.path.exists(src):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    agg = {}
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v_us
    ours = {k: v for k, v in agg.items() if "nplda" in k or "::" in k and "at::" not in k}
    tot = sum(v[1] for v in ours.values())
    with open(os.path.join(P, f"{rnd}_train_launches.csv"), "w") as f:
        f.write(f"# ncu launch list of training steps, {rnd} (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)\n")
        f.write("# command: ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/ncu_train.py 1000000 d\n")
        f.write("# one NeuralPlda step and one DPlda step (LDA frozen) on 1M pairs: forward, BCE, backward; libnplda kernels only\n")
        f.write("share_pct,launches,avg_us,kernel\n")
        for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{100 * v[1] / tot:.2f},{v[0]},{v[1] / v[0]:.1f},\"{k[:110]}\"\n")
    print("wrote train launches", len(ours))

# ---- full captures: every gpurun_out/<tag>_<name>.ncu-rep -> profiles/<rnd>_<name>_ncu.csv (first launch in the report)
def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]

def git_head():
    return subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()

for fn in sorted(os.listdir(G)):
    if not (fn.startswith(tag + "_") and fn.endswith(".ncu-rep")):
        continue
    name = fn[len(tag) + 1:-len(".ncu-rep")]
    outname = f"{rnd}_{name}_ncu.csv"
    hdr, units, launches = raw(os.path.join(G, fn))
    vals = launches[0]
    kname = vals[hdr.index("Kernel Name")]
    with open(os.path.join(P, outname), "w") as f:
        f.write(f"# ncu capture, {rnd} (commit {git_head()}; exact command in tools/r2_profiles.sh): {kname[:120]}, B200\n")
        f.write("metric,unit,value\n")
        got = {}
        for h, u, v in zip(hdr, units, vals):
            base = h.split("TriageCompute.")[-1]
            if any(base == k for k in KEYS):
                f.write(f"{base},{u},{v}\n"); got[base] = (u, v)
    print("wrote", outname)
    if name == "score_tcp" and "dram__bytes_read.sum" in got:       # the kernel the bench step launches (K1p)
        def tobytes(u, v):
            v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        rd, wr = tobytes(*got["dram__bytes_read.sum"]), tobytes(*got["dram__bytes_write.sum"])
        json.dump({"kernel": kname[:80], "pairs": 1000000, "dram_bytes_read": rd, "dram_bytes_write": wr, "commit": git_head(),
                   "source": f"ncu --set full --clock-control none, one launch of `python bench.py --steps 1 --warmup 3 --no-cpu-baseline "
                             f"--skip-e2e` (tools/r2_profiles_final.sh); summary in profiles/{outname}"},
                  open(os.path.join(P, "k1_traffic.json"), "w"), indent=1)
        print("wrote k1_traffic.json", rd, wr)
