"""In-graph time attribution by ablation: replays the full-arch forward with one kernel class removed at a time (RDM_SKIP bit mask,
csrc/unet.cu) and prints how much the forward shrinks.  python tools/ablate_forward.py [mode]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mode = sys.argv[1] if len(sys.argv) > 1 else "3"
names = {0: "full forward", 1: "gn_stats", 2: "gn_apply", 4: "layernorm", 8: "attention", 16: "GEMM M>=8192", 32: "GEMM M<8192", 48: "all GEMMs", 15: "all glue"}
base = None
for bit, nm in names.items():
    env = dict(os.environ, RDM_SKIP=str(bit))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_forward.py"), mode, "30"], env=env, capture_output=True, text=True).stdout
    ms = float(out.split("graph forward:")[1].split()[0])
    base = ms if base is None else base
    print(f"skip {nm:14s}: {ms:7.3f} ms   (-{base - ms:.3f})", flush=True)
