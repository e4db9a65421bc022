"""Turn the raw outputs of scripts/gpu_round.sh (gpurun_out/<tag>_*) into the committed evidence under profiles/:
   <tag>_bench.json             the bench line of the run
   <tag>_launches.md            per-kernel launch count / time / share of ONE sample() step (ncu launch list, serialised, cold cache)
   <tag>_ncu_top_kernels.md     raw-page excerpts of the ncu --set full capture (first instance of each kernel)
   traffic.json                 DRAM bytes per launch of the kernel classes bench.py reports a roofline for (from the same capture)
   <tag>_sass.md                UTC*MMA / LDTM / STTM / UTMALDG / UTMASTG / HMMA counts per kernel of the shipped liblamslide.so
Usage: python scripts/collect_profiles.py r02"""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()

# ---- bench line
bench = json.loads(open(os.path.join(G, f"{tag}_bench.json")).read().strip().splitlines()[-1])
json.dump(bench, open(os.path.join(P, f"{tag}_bench.json"), "w"), indent=1)
per_step = int(round(bench["gpu_launches"] / bench["steps"]))

# ---- launch list of one step
rows = []
with open(os.path.join(G, f"{tag}_launches.csv")) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        rows.append((r["Kernel Name"], v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r["Metric Unit"], 1.0)))
# the launch list holds torch's kernels too (noise, copies): take the window that ends with the last launch and starts at the first
# first-stage kernel of the last step
names = [n for n, _ in rows]
last_step_start = max(i for i, n in enumerate(names) if "gather_cols" in n and i < len(names) - 50 and
                      not any("gather_cols" in m for m in names[max(0, i - 3):i]))
sel = rows[last_step_start:]
agg = OrderedDict()
for name, us in sel:
    short = re.sub(r"^void (lam::)?", "", re.sub(r"\(.*", "", name))
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, f"{tag}_launches.md"), "w") as f:
    f.write(f"# {tag} — ncu launch list of one `sample()` step (4AA, B=64, 1xB200), commit {head}\n\n")
    f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline "
            "--no-profile --no-secondary` (scripts/gpu_round.sh).  Per-launch times are serialised and cold-cache: compare SHARES with the "
            f"live CUDA-event shares of `{tag}_bench.json`, not absolutes.\n\n")
    f.write(f"launches in window: {len(sel)} (bench counts {per_step} launches of this library per step), total device time {tot / 1e3:.2f} ms\n\n")
    f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {c} | {us / 1e3:.3f} | {100 * us / tot:.1f}% | {us / c:.1f} |\n")
    f.write("\nLive CUDA-event shares of the un-profiled bench run (`kernel_time_shares`): " +
            ", ".join(f"{k} {100 * v:.1f} %" for k, v in sorted(bench["kernel_time_shares"].items(), key=lambda kv: -kv[1])) + "\n")

# ---- ncu --set full excerpts + traffic
rep = os.path.join(G, f"{tag}_top.ncu-rep")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(out)))
hdr, units = rr[0], rr[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__cycles_elapsed.max', 'smsp__sass_inst_executed_op_tmem_ldt.sum', 'smsp__sass_inst_executed_op_tmem_stt.sum']
seen, traffic = OrderedDict(), {}
classes = {"attn_tc_kernel": "attn_temporal", "mlp_fused_kernel": "gemm_linear2", "ln_modulate_kernel<3, 0>": "ln_modulate",
           "EpiLinear1Ws<24, 0>": "gemm_linear1", "EpiLinear1Ws<24, 2>": "gemm_linear1_spatial"}
with open(os.path.join(P, f"{tag}_ncu_top_kernels.md"), "w") as f:
    f.write(f"# {tag} — ncu --set full of the top kernels (4AA, B=64: 128k tokens per launch), raw-page excerpts, commit {head}\n\n```\n")
    for r in rr[2:]:
        full = r[idx['Kernel Name']]
        name = re.sub(r"^void (lam::)?", "", full.split('(')[0])
        if name in seen:
            continue
        seen[name] = 1
        f.write(name + "\n")
        for w in want:
            if w in idx:
                f.write(f"    {w:82s} {r[idx[w]]:>18s} {units[idx[w]]}\n")

        def num(k):
            v, u = float(r[idx[k]].replace(",", "")), units[idx[k]]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
        for pat, cls in classes.items():
            if pat in name and cls not in traffic:
                traffic[cls] = {"dram_bytes_per_launch": int(num('dram__bytes_read.sum') + num('dram__bytes_write.sum')),
                                "ncu_us_per_launch": round(num('gpu__time_duration.sum'), 3), "kernel": name}
    f.write("```\n")
json.dump({"source": f"ncu --set full --clock-control none, profiles/{tag}_ncu_top_kernels.md, commit {head} (4AA, B=64, 1xB200): "
                     "dram__bytes_read.sum + dram__bytes_write.sum per launch", "kernels": traffic},
          open(os.path.join(P, "traffic.json"), "w"), indent=1)

# ---- SASS evidence of the shipped library
so = os.path.join(ROOT, "lam_slide_b200", "lib", "liblamslide.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, counts = None, OrderedDict()
pats = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA", "MUFU.EX2", "FFMA2"]
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"^void (lam::)?", "", cur.split('(')[0])
        counts[cur] = dict.fromkeys(pats, 0)
        continue
    if cur:
        for p_ in pats:
            if re.search(r"\b" + re.escape(p_), line):
                counts[cur][p_] += 1
with open(os.path.join(P, f"{tag}_sass.md"), "w") as f:
    f.write(f"# {tag} — SASS mnemonic counts per kernel of the shipped `liblamslide.so` (`cuobjdump -sass`), commit {head}\n\n")
    f.write("tcgen05.mma -> UTCHMMA, tcgen05.ld / st -> LDTM / STTM, TMA -> UTMALDG / UTMASTG / UTMAREDG, mma.sync -> HMMA.\n\n")
    f.write("| kernel | " + " | ".join(pats) + " |\n|---|" + "---:|" * len(pats) + "\n")
    for k, c in counts.items():
        if any(c[p_] for p_ in pats[:7]):
            f.write(f"| `{k}` | " + " | ".join(str(c[p_]) for p_ in pats) + " |\n")
print("wrote profiles for", tag, "at", head)
