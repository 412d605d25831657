#!/usr/bin/env python
"""CPU-side: turn what a gpurun profiling call left in gpurun_out/ into the tracked
evidence under profiles/ (ncu raw-page summaries per kernel, launch list, SASS mnemonic
census + listings of our kernels, kernel microbench table)."""
import collections
import csv
import glob
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_active.avg", "lts__t_sector_hit_rate.pct",
]


def ncu_raw(rep):
    p = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(p.stdout.splitlines()))
    if len(rows) < 3:
        return [], []
    return rows[0], rows[2:]


def summarize_rep(rep, title):
    hdr, rows = ncu_raw(rep)
    if not hdr:
        return None
    idx = {m: hdr.index(m) for m in METRICS if m in hdr}
    name_i = hdr.index("Kernel Name")
    lines = [f"# {title}", "", f"source: `{os.path.relpath(rep, ROOT)}` (ncu --set full --clock-control none, cold L2 per replay)", "",
             "| # | kernel | " + " | ".join(m.split(".")[0].replace("__", " ") for m in idx) + " |",
             "|---|---|" + "---|" * len(idx)]
    for i, r in enumerate(rows):
        nm = re.sub(r"\(.*", "", r[name_i]).replace("void ssb::", "")
        lines.append(f"| {i} | `{nm}` | " + " | ".join(r[j] for j in idx.values()) + " |")
    return "\n".join(lines) + "\n"


def launches_md(path, title):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    out = [f"# {title}", "", "ncu --metrics gpu__time_duration.sum (serialised, cold cache: compare SHARES, not absolutes)", "",
           "| id | kernel | grid | stream | us |", "|---|---|---|---|---|"]
    tot = 0.0
    agg = collections.defaultdict(float)
    for r in rows:
        v = float(r["Metric Value"].replace(",", "")) / 1e3
        tot += v
        nm = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        agg[nm] += v
        out.append(f"| {r['ID']} | `{nm}` | {r['Grid Size']} | {r['Stream']} | {v:.2f} |")
    out += ["", f"sum = {tot:.1f} us over {len(rows)} launches", "", "| kernel | total us | share |", "|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        out.append(f"| `{k}` | {v:.1f} | {100 * v / tot:.1f}% |")
    return "\n".join(out) + "\n"


def sass_census():
    so = glob.glob(os.path.join(ROOT, "shallowspeed_b200", "_C*.so"))
    if not so:
        return
    p = subprocess.run(["cuobjdump", "-sass", so[0]], capture_output=True, text=True)
    text = p.stdout
    funcs = re.split(r"\n\s*Function : ", text)
    census = ["# SASS evidence (cuobjdump -sass of the in-tree extension, sm_100a)", "",
              "| kernel | UTCHMMA (tcgen05.mma) | LDTM (tcgen05.ld) | UTMALDG (TMA load) | UTMASTG/UTMAREDG (TMA store/reduce) | "
              "UTCBAR (tcgen05.commit) | MEMBAR.*SYS | ST/LD .SYS (peer flags) | HMMA (legacy) | .MULTICAST (cluster TMA / commit) | LDGMC (multimem.ld_reduce) |",
              "|---|---|---|---|---|---|---|---|---|---|---|"]
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        if "ssb" not in name:
            continue
        c = lambda pat: len(re.findall(pat, f))
        census.append(f"| `{name[:70]}` | {c(r'UTCHMMA')} | {c(r'LDTM')} | {c(r'UTMALDG')} | {c(r'UTMASTG') + c(r'UTMAREDG')} | "
                      f"{c(r'UTCBAR')} | {c(r'MEMBAR\.[A-Z.]*SYS')} | {c(r'\.SYS') - c(r'MEMBAR\.[A-Z.]*SYS')} | {c(r'HMMA') - c(r'UTCHMMA')} | "
                      f"{c(r'\.MULTICAST')} | {c(r'LDGMC')} |")
        if any(k in name for k in ("tc_gemm_kernel", "fused_wgrad_dp", "mlp_chain_kernel", "dp_ll_wgrad", "tc_wgrad_group", "nvls_reduce_sgd")):
            if "fused_wgrad_dp" in name:
                short = "fused_wgrad_dp"
            elif "mlp_chain_kernel" in name:
                m3 = re.search(r"ILb(\d)ELb(\d)ELb(\d)E", name)
                short = "mlp_chain_" + ("fp32" if m3.group(1) == "1" else "tf32") + ("_fold" if m3.group(2) == "1" else "") + ("_acc" if m3.group(3) == "1" else "")
            elif "dp_ll_wgrad" in name:
                short = "dp_ll_dp" + re.search(r"ILi(\d)E", name).group(1)
            elif "tc_wgrad_group" in name:
                short = "tc_wgrad_group"
            elif "nvls_reduce_sgd" in name:
                short = "nvls_reduce_sgd"
            else:
                m = re.search(r"ILi(\d)E(?:Lb(\d)E)?", name)
                short = "tc_gemm_mode" + m.group(1) + ("_splitk" if m.group(2) == "1" else "")
            keep = [ln for ln in f.split("\n") if re.search(r"UTC|LDTM|UTMA|SYNCS|MEMBAR|\.SYS|UBLKCP|ELECT|BAR\.|\.STRONG\.GPU|RED\.|LDGMC|MULTIMEM|FENCE", ln)]
            open(os.path.join(OUT, f"sass_{short}.txt"), "w").write(
                f"// {name}\n// tensor-core / TMA / barrier / system-scope instructions only (full listing: cuobjdump -sass)\n" + "\n".join(keep) + "\n")
    open(os.path.join(OUT, "sass_census.md"), "w").write("\n".join(census) + "\n")


def kernel_bench_md():
    path = os.path.join(GO, "kernel_bench.jsonl")
    if not os.path.exists(path):
        return
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    out = ["# GEMM family microbenchmark (CUDA events, L2 flushed between timed launches, median)", "",
           f"roofline denominators: HBM copy {peaks.get('hbm_gbs', 6650)} GB/s (of measured), bf16 {peaks.get('bf16_tflops', 1590)} TFLOP/s", "",
           "`cold`: L2 flushed by a read-modify-write pass before every launch (leaves dirty lines: pessimistic).  `stream`: back-to-back launches "
           "rotating over >= 3 x L2 of distinct weight buffers (what a step of a model larger than L2 looks like).", "",
           "| kernel | rows | in | out | us (cold L2) | GB/s | frac of measured HBM | TFLOP/s (tf32) | us (stream) | GB/s (stream) | frac of measured HBM (stream) |",
           "|---|---|---|---|---|---|---|---|---|---|---|"]
    for l in open(path):
        try:
            d = json.loads(l)
        except Exception:
            continue
        out.append(f"| {d['kernel']} | {d['rows']} | {d['in']} | {d['out']} | {d['us_cold_l2']} | {d['GBps_cold']} | {d['frac_hbm_measured']} | {d['tflops_cold']} | "
                   f"{d.get('us_stream', '')} | {d.get('GBps_stream', '')} | {d.get('frac_hbm_stream', '')} |")
    open(os.path.join(OUT, "kernel_bench.md"), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    for rep in glob.glob(os.path.join(GO, "prof_*.ncu-rep")):
        tag = os.path.basename(rep)[5:-8]
        md = summarize_rep(rep, f"ncu full capture: {tag}")
        if md:
            open(os.path.join(OUT, f"ncu_{tag}.md"), "w").write(md)
    for path in glob.glob(os.path.join(GO, "launches_*.csv")):
        tag = os.path.basename(path)[9:-4]
        open(os.path.join(OUT, f"launches_{tag}.md"), "w").write(launches_md(path, f"launch list: {tag}"))
    kernel_bench_md()
    sass_census()
    for f in glob.glob(os.path.join(GO, "bench_*.json")):
        txt = [l for l in open(f) if l.startswith("{")]
        if txt:
            open(os.path.join(OUT, os.path.basename(f)), "w").write(txt[-1])
    print("profiles/:", sorted(os.listdir(OUT)))
