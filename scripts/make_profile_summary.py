"""Builds the committed ncu evidence under profiles/ from the scratch captures in gpurun_out/ (run here, no GPU):
   python scripts/make_profile_summary.py r1
Writes profiles/<round>_launches_step.csv (raw ncu launch list of one PPO iteration), profiles/<round>_ncu_summary.md."""
import csv, io, os, re, shutil, subprocess, sys, collections, json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.environ.get("CRUX_PROFILE_SRC", os.path.join(ROOT, "gpurun_out"))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = []


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot, n = collections.OrderedDict(), 0.0, 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")[:48]
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] == "ns" else v * 1000 if row["Metric Unit"] == "ms" else v
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; tot += v; n += 1
    return agg, tot, n


def metrics(rep, names):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = {}
    for w in names:
        if w in hdr:
            i = hdr.index(w)
            res[w] = (rows[2][i], units[i])
    res["kernel"] = rows[2][hdr.index("Kernel Name")][:70]
    return res


KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]

lp = os.path.join(G, "launches_step.csv")
if os.path.exists(lp):
    shutil.copy(lp, os.path.join(ROOT, "profiles", f"{rnd}_launches_step.csv"))
    agg, tot, n = launches(lp)
    out += [f"## One PPO iteration (BASELINE config[1], device env) — ncu launch list", "",
            f"`ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python scripts/ncu_step.py` "
            f"(raw: `profiles/{rnd}_launches_step.csv`).  {n} launches, {tot:.0f} us summed (cold-cache, serialised: compare SHARES).", "",
            "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| `{k}` | {c} | {v:.1f} | {v / c:.2f} | {100 * v / tot:.1f}% |")
    out.append("")
lp2 = os.path.join(G, "launches_sac.csv")
if os.path.exists(lp2):
    shutil.copy(lp2, os.path.join(ROOT, "profiles", f"{rnd}_launches_sac.csv"))
    agg, tot, n = launches(lp2)
    out += [f"## One SAC update (BASELINE config[3]: 376-obs / 17-act, 256-256, batch 2048) — ncu launch list", "",
            f"`ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python scripts/sac_launches.py` "
            f"(raw: `profiles/{rnd}_launches_sac.csv`).  {n} launches, {tot:.0f} us summed (cold-cache, serialised; live: `scripts/sac_launches.py --time`).", "",
            "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| `{k}` | {c} | {v:.1f} | {v / c:.2f} | {100 * v / tot:.1f}% |")
    out.append("")
for name, title in (("r2_prof_minibatch", "mb6::minibatch_kernel<0> (actor; dominant kernel of the step): every GEMM on tcgen05, TMEM accumulators"), ("prof_minibatch", "fused_minibatch_kernel (dominant kernel of the step)"), ("prof_gae_tma", "gae_tma_kernel at [2048, 16384] (738 MB): TMA-fed streaming scan, the kernel crux_fill_gae_returns runs for N >= 8192"),
                    ("prof_gae", "gae_returns_kernel at [2048, 16384] (738 MB): register-resident scan (narrow / short rollouts; CRUX_GAE=scan)"),
                    ("prof_fwd_tc5", "tc5::forward_kernel_tmem: value(V, s) over 131072 rows on tcgen05, activations resident in tensor memory"),
                    ("prof_mb5", "mb5::minibatch_kernel (opt-in CRUX_MB_TC5=1): PPO minibatch with the row GEMMs on tcgen05 / TMEM"),
                    ("r2_prof_tail", "reduce_adam_kernel: the single-launch update tail (partials -> gradient -> [LL exchange] -> Adam -> planes -> record), 187 CTAs of 128 threads"),
                    ("r2_prof_gemm_tc5", "gemm_tc5_kernel: the layer engine's Dense GEMM on tcgen05 (SAC 376/17/256-256, B = 2048: forward 2048 x 256 x 393; A operand in tensor memory, 3xTF32)"),
                    ("r2_prof_rollout", "rollout_linquad_kernel (r2 capture)"), ("prof_rollout", "rollout_linquad_kernel: T = 32 vector steps of 4096 env streams in one persistent launch"),
                    ("prof_persist", "mbp::epoch_kernel (persistent 8-CTA cluster kernel, 256 minibatches of 128 rows in one launch)"),
                    ("prof_forward", "fused_forward_kernel")):
    rep = os.path.join(G, name + ".ncu-rep")
    if not os.path.exists(rep):
        continue
    m = metrics(rep, KEYS)
    out += [f"## {title} — `ncu --set full --clock-control none --import-source on` (first captured launch)", "", f"kernel: `{m['kernel']}`", "",
            "| metric | value | unit |", "|---|---:|---|"]
    for k in KEYS:
        if k in m:
            out.append(f"| {k} | {m[k][0]} | {m[k][1]} |")
    out.append("")
# dram traffic per launch of the captured kernels (bench.py reports it as roofline.traffic)
traffic = {}
for name, key in (("prof_minibatch", "fused_minibatch_kernel"), ("prof_gae", "gae_returns_kernel"), ("prof_gae_tma", "gae_tma_kernel")):
    rep = os.path.join(G, name + ".ncu-rep")
    if os.path.exists(rep):
        m = metrics(rep, ["dram__bytes_read.sum", "dram__bytes_write.sum"])
        def to_bytes(v, u):
            v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        traffic[key] = sum(to_bytes(*m[k]) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in m)
tpath = os.path.join(ROOT, "profiles", f"{rnd}_traffic.json")
if os.path.exists(tpath):          # hand-annotated entries (cold / warm L2 captures, notes) are kept; captured kernels are refreshed
    old = json.load(open(tpath))
    old.update(traffic)
    traffic = old
json.dump(traffic, open(tpath, "w"), indent=1)
path = os.path.join(ROOT, "profiles", f"{rnd}_ncu_summary.md")
hdr = [f"# ncu evidence, {rnd} (generated by scripts/make_profile_summary.py from gpurun_out/ captures on a B200)", ""]
notes = os.path.join(ROOT, "profiles", f"{rnd}_notes.md")
open(path, "w").write("\n".join(hdr + out) + "\n")
print(path)
