"""Turn the raw ncu outputs of tools/prof_round.sh (gpurun_out/) into the tracked summaries under profiles/.
usage: python tools/summarize_profiles.py <launches.csv> <full.ncu-rep> [round tag]"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches, rep = sys.argv[1], sys.argv[2]
tag = sys.argv[3] if len(sys.argv) > 3 else "r1"

rows = list(csv.reader(open(launches)))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
ki, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, tot, n = collections.OrderedDict(), 0.0, 0
for r in rows[hi + 1:]:
    if len(r) <= mi:
        continue
    v, u = float(r[mi].replace(",", "")), r[ui]
    ms = v / 1e6 if u.startswith("ns") else v / 1e3 if u.startswith("us") else v if u.startswith("ms") else v * 1e3
    short = re.sub(r"\(.*", "", r[ki])[:100]
    d = agg.setdefault(short, [0.0, 0])
    d[0] += ms
    d[1] += 1
    tot += ms
    n += 1
lx = sum(v[0] for k, v in agg.items() if "lx::" in k)
out = [
    f"{tag} (final kernels of this round) — ncu launch list (gpu__time_duration.sum, --clock-control none) of",
    "  python bench.py --layers 2 --steps 1 --warmup 3 --no-cpu-baseline   (6 optimizer steps: 3 warm-up + 1 timed + 1 instrumented + 1 e2e)",
    "per-launch times are cold-cache and serialised: compare SHARES, not absolutes. With 2 of 32 layers the",
    "per-step-constant LM head (8 bf16 GEMM launches + cross-entropy) and optimizer weigh 16x more than in the real step;",
    f"the 32-layer shares measured live by bench.py (CUDA events) are in profiles/{tag}_bench_1gpu.json and DESIGN.md.",
    f"total {tot:.1f} ms over {n} launches; llamax_b200 kernels (lx::*) = {100 * lx / tot:.1f} % of device time",
    "",
]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    out.append(f"{100 * v[0] / tot:6.2f}% {v[0]:10.2f} ms  n={v[1]:5d}  {k}")
open(os.path.join(ROOT, "profiles", f"{tag}_launch_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:24]))

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
want = ["ID", "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "SM_A.TriageCompute.sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg.per_second", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum"]
idx = [h.index(w) for w in want if w in h]
tab = [[h[i] for i in idx], [units[i] for i in idx]]
for r in rr[2:]:
    row = [r[i] for i in idx]
    row[1] = row[1][:90]
    tab.append(row)
csv.writer(open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_hot_kernels.csv"), "w")).writerows(tab)
for r in tab:
    print(r[:9])


# dram traffic of the dominant kernel for bench.py's roofline.traffic (keyed by kernel + shape; tools/prof_kernels.py
# launches the bf16 grad_input GEMM [M=16384, N=4096, K=28688] first)
import json

git = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
traffic = {}
hn = tab[0]
for r in tab[2:]:
    name = r[hn.index("Kernel Name")]
    if "gemm_wide_kernel" in name or name.startswith("void gemm_kernel<0, 2, 0, 0, 0, 0>") or "lx::gemm_kernel<0, 2, 0, 0, 0, 0>" in name:
        rd, wr = float(r[hn.index("dram__bytes_read.sum")]), float(r[hn.index("dram__bytes_write.sum")])
        unit_r, unit_w = tab[1][hn.index("dram__bytes_read.sum")], tab[1][hn.index("dram__bytes_write.sum")]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        nbytes = rd * scale[unit_r] + wr * scale[unit_w]
        kname = "gemm_wide_kernel<bf16,cta_group::2,512x256>" if "wide" in name else "gemm_kernel<bf16,cta_group::2>"
        traffic[kname + "[M=16384,N=4096,K=28688]"] = {
            "bytes": int(nbytes), "git": git, "source": f"profiles/{tag}_ncu_full_hot_kernels.csv ID {r[0]} (dram__bytes_read.sum + dram__bytes_write.sum, one launch, ncu --set full)",
            "algorithmic_bytes": int(2 * (16384 * 28688 + 4096 * 28688 + 16384 * 4096)),
            "l2_hit_pct": float(r[hn.index("lts__t_sector_hit_rate.pct")])}
        break
json.dump(traffic, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(traffic)
