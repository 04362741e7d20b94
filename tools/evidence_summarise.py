"""Turns the ncu / sanitizer outputs of tools/evidence.sh (gpurun_out/, scratch) into the tracked summaries under profiles/.
python tools/evidence_summarise.py r2"""
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}
METRICS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]


def short(name):
    return re.sub(r"void |<unnamed>::|\(anonymous namespace\)::", "", name).split("(")[0][:64]


GEMMS_PER_FWD = 182      # tcgen05 GEMM launches of one full-architecture forward (197 before the skip_connection K extension)


def gemm_traffic():
    src = os.path.join(G, f"gemm_tc_dram_{tag}.csv")
    rows = list(csv.DictReader([l for l in open(src) if not l.startswith("==")]))
    tot, ids = collections.Counter(), set()
    for x in rows:
        tot[x["Metric Name"]] += float(x["Metric Value"].replace(",", "")) * SCALE.get(x["Metric Unit"], 1)
        ids.add(x["ID"])
    n = len(ids)
    per_fwd = (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) * GEMMS_PER_FWD / n
    shutil.copy(src, os.path.join(P, f"gemm_tc_dram_{tag}.csv"))
    out = {"bytes_per_forward": per_fwd, "launches_captured": n, "dram_read_bytes": tot["dram__bytes_read.sum"], "dram_write_bytes": tot["dram__bytes_write.sum"],
           "sum_gpu_time_ms": tot["gpu__time_duration.sum"] / 1e6,
           "how": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:gemm_tc_kernel over {n} of the {GEMMS_PER_FWD} tcgen05 GEMM launches of one full-architecture forward "
                  f"(B2 = 32, fp16 mode), scaled by {GEMMS_PER_FWD}/{n}; profiles/gemm_tc_dram_{tag}.csv; every launch is replayed with a flushed L2, so activations that are L2 hits in "
                  "the running step count as DRAM reads here (an upper bound of the in-step traffic)"}
    json.dump(out, open(os.path.join(P, f"gemm_tc_dram_{tag}.json"), "w"), indent=1)
    print("traffic", out["bytes_per_forward"] / 1e9, "GB per forward from", n, "launches")


def launches(name, title, skip):
    src = os.path.join(G, f"launches_{name}_{tag}.csv")
    shutil.copy(src, os.path.join(P, f"launches_{name}_{tag}.csv"))
    rows = list(csv.DictReader([l for l in open(src) if not l.startswith("==")]))
    rows = rows[int(len(rows) * skip):]
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0]); a[0] += 1; a[1] += float(r["Metric Value"].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, f"launches_{name}_{tag}.md"), "w") as f:
        f.write(f"# {title}\n\n{len(rows)} launches, {tot / 1e6:.3f} ms total (cold-cache, serialised under ncu: compare SHARES)\n\n| kernel | launches | us | share | us / launch |\n|---|---|---|---|---|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {c} | {t / 1e3:.1f} | {100 * t / tot:.1f}% | {t / 1e3 / c:.2f} |\n")


def raw_table(rep, title, how, f, pick=None, stalls=True):
    if not os.path.exists(os.path.join(G, rep)):
        f.write(f"\n## {title}\n\n(no capture in this run: `{rep}` missing)\n")
        return
    out = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    if pick:
        data = [data[i] for i in pick if i < len(data)]
    ki = hdr.index("Kernel Name")
    f.write(f"\n## {title}\n\n`{how}`\n\n| metric | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |\n|---|" + "---|" * len(data) + "\n")
    f.write("| kernel | " + " | ".join(f"`{short(d[ki])}`" for d in data) + " |\n")
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            f.write(f"| `{m}` [{units[i]}] | " + " | ".join(d[i] for d in data) + " |\n")
    if stalls:
        st = [i for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
        top = sorted(st, key=lambda i: -max(float(d[i] or 0) for d in data))[:6]
        for i in top:
            f.write(f"| stall `{hdr[i][len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]}` [warps / issue] | " + " | ".join(f"{float(d[i] or 0):.2f}" for d in data) + " |\n")


gemm_traffic()
launches("bench", "ncu launch list: `ncu --metrics gpu__time_duration.sum --clock-control none -c 3000` over `python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-r-shape` "
         "(database build, kNN, context K/V, the first DDIM steps of the fp16 engine; last 80 % of the list)", 0.2)
if os.path.exists(os.path.join(G, f"launches_rarm_{tag}.csv")):
    launches("rarm", "ncu launch list of the RARM decode loop: `ncu --metrics gpu__time_duration.sum -k regex:rarm_ --launch-skip 600 -c 300` over `python tools/rarm_bench.py` "
             "(two decode steps of the ImageNet-size model, batch 4, fp16 weights)", 0.0)
with open(os.path.join(P, f"ncu_summary_{tag}.md"), "w") as f:
    f.write(f"# ncu `--set full --clock-control none --import-source on` captures, round 2 (B200, sm_100a)\n\nRaw reports stay in `gpurun_out/` (scratch); the tables are "
            "`ncu -i <rep> --page raw --csv` of those files (tools/evidence.sh, tools/evidence_summarise.py).  Durations are cold-cache, serialised ncu replays.\n")
    if os.path.exists(os.path.join(G, "gemm_tc_r2g.ncu-rep")):
        raw_table("gemm_tc_r2g.ncu-rep", "gemm_tc_kernel (fp16 mode, one MMA per product): 12 consecutive launches of a graph-replayed full-architecture forward (B2 = 32)",
                  "ncu --set full -k regex:gemm_tc_kernel --launch-skip 203 --launch-count 12 python tools/profile_forward.py 4 1  [before the epilogue rework of this round: "
                  "launch 0 = level-0 3x3 conv M = 32768; 1-9 level-1 layers incl. the fused cross-attention (8) ; 10 = GEGLU feed-forward]", f)
    raw_table(f"glue_{tag}.ncu-rep", "glue kernels of the forward: GroupNorm statistics / apply (large images), one-launch cluster GroupNorm (small images), LayerNorm, warp-MMA self-attention",
              "ncu --set full -k regex:'gn_stats|gn_apply|gn_fused|layernorm_kernel|attention_mma' --launch-skip 150 --launch-count 10 python tools/profile_forward.py 4 1", f)
    raw_table(f"rarm_{tag}.ncu-rep", "RARM decode step: weight-streaming GEMV (`rarm_gemv_kernel`) and cached attention (`rarm_attn_kernel`)",
              "ncu --set full -k regex:'rarm_gemv|rarm_attn' --launch-skip 600 --launch-count 9 python tools/rarm_bench.py", f)
    raw_table(f"knn_{tag}.ncu-rep", "one exact kNN search, fp16 database 1,281,167 x 512, 16 queries, k = 4 (normalise + fused tensor-core scan + select + conditional fallback pair)",
              "ncu --set full -k regex:knn_ --launch-skip 12 --launch-count 6 python tools/knn_sweep.py --n 1281167 --q 16 --dtypes float16", f)
    # later captures of the round (present only for some tags)
    if os.path.exists(os.path.join(G, f"gemm_{tag}.ncu-rep")):
        raw_table(f"gemm_{tag}.ncu-rep", "gemm_tc_kernel, final state of the round (skip_connection K extension, upsample fold): 12 consecutive launches of a graph-replayed forward",
                  "ncu --set full -k regex:gemm_tc_kernel --launch-skip 188 --launch-count 12 python tools/profile_forward.py 4 1", f)
    if os.path.exists(os.path.join(G, f"convfirst_{tag}.ncu-rep")):
        raw_table(f"convfirst_{tag}.ncu-rep", "weight-load kernels and the register-tiled first convolution (`fold_up_weights_kernel`, `split_planes_kernel`, `conv_first_kernel`)",
                  "ncu --set full -k regex:'conv_first|split_planes|fold_up' --launch-count 3 python tools/profile_forward.py 4 1", f)
    if os.path.exists(os.path.join(G, f"knn64_{tag}.ncu-rep")):
        raw_table(f"knn64_{tag}.ncu-rep", "fused tensor-core kNN scan at 64 queries (hi-only pass), 1,281,167 x 512 fp16: 1.317 GB in 226.6 us = 5.81 TB/s = 0.90 of the measured copy peak",
                  "ncu --set full -k regex:knn_scan_fused --launch-skip 4 --launch-count 1 python tools/knn_sweep.py --n 1281167 --q 64 --k 4 --dtypes float16", f)
    if os.path.exists(os.path.join(G, f"knnsel_{tag}.ncu-rep")):
        raw_table(f"knnsel_{tag}.ncu-rep", "knn_select_kernel (16 queries): 21.7 us; 44 % of the stall samples sit at the barrier behind the exact re-rank (one lane per candidate "
                  "evaluates the oracle's SEQUENTIAL 512-term fp64 sum: ~35 cycles per dependent DADD)",
                  "ncu --set full -k regex:knn_select --launch-skip 8 --launch-count 1 python tools/knn_sweep.py --n 1281167 --q 16 --k 4 --dtypes float16", f)
if os.path.exists(os.path.join(G, f"sanitize_{tag}.log")):
    shutil.copy(os.path.join(G, f"sanitize_{tag}.log"), os.path.join(P, f"sanitize_{tag}.log"))
with open(os.path.join(P, f"sass_histogram_{tag}.md"), "w") as f:
    f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_histogram.py")], capture_output=True, text=True).stdout)
print("done")
