"""Fills the @@PLACEHOLDER@@ fields of DESIGN.md's round-2 table from the snapshot files under profiles/ (tools/snapshot.sh <tag>).
python tools/fill_design.py <bench tag> <knn tag>   (idempotent once filled: placeholders are gone)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, ktag = sys.argv[1], sys.argv[2]
P = lambda f: os.path.join(ROOT, "profiles", f)
j = json.load(open(P(f"bench_{tag}.json")))
x2 = json.load(open(P(f"bench_{tag}_fp16x2.json")))
r, rs = j["roofline"], j["r_shape"]
def sweep(f):
    return [json.loads(l) for l in open(P(f))] if f.endswith(".jsonl") else json.load(open(P(f)))
def rows(f):
    d = json.load(open(P(f)))
    return d if isinstance(d, list) else d["rows"]
k1 = rows(f"knn_sweep_{ktag}_1m_k4.json")
k20 = rows(f"knn_sweep_{ktag}_20m_k4.json")
k50 = rows(f"knn_sweep_{ktag}_50m_k4.json")
f16 = [x for x in k1 if x["dtype"] == "float16"]
knn_tab = ", ".join(f"{x['nq']} q: {x['frac_hbm']:.2f}" for x in f16) + f" of the measured 6.46 TB/s copy peak ({f16[-1]['qps'] / 1e3:.0f} K queries/s at {f16[-1]['nq']} queries)"
knn_big = "; ".join(f"{x['n'] / 1e6:.1f} M rows: " + ", ".join(f"{y['nq']} q {y['frac_hbm']:.2f}" for y in grp) for x, grp in ((k20[0], k20), (k50[0], k50)))
rep = {"VALUE": f"{j['value']:.1f}", "E2E": f"{j['e2e']['value']:.1f}", "FP16X2": f"{x2['value']:.1f}", "TAG": tag, "KTAG": ktag,
       "RSHAPE": f"{rs['value']:.1f}", "RSHAPE_E2E": f"{rs['e2e']:.1f}", "RSHAPE_FWD": f"{rs['forward_ms_graph']:.1f}", "RSHAPE_FRAC": f"{rs['frac_of_sustained_peak_whole_step']:.2f}",
       "FWD": f"{r['forward_ms_graph']:.2f}", "TCMS": f"{r['tc_ms_per_forward']:.2f}", "GLUE": f"{r['forward_ms_graph_without_gemms']:.2f}",
       "ACH": f"{r['achieved']:.0f}", "FRAC": f"{r['frac']:.3f}", "WFRAC": f"{r['whole_step_frac']:.2f}", "KNNTABLE": knn_tab, "KNNBIG": knn_big,
       "KNNFRAC": f"{j['knn']['frac_hbm']:.2f}"}
s = open(os.path.join(ROOT, "DESIGN.md")).read()
for k, v in rep.items():
    s = s.replace(f"@@{k}@@", v)
open(os.path.join(ROOT, "DESIGN.md"), "w").write(s)
print(rep)
