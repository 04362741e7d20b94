"""Builds tests/emu/_build/librdm_emu.so: the SIMT translation units of csrc/ (rarm.cu, unet.cu, kernels.cu, gemm_simt.cu, common.cu) compiled by g++ against the host emulation of CUDA in
tests/emu/include (see cuda_runtime.h there).  The sources are used as written except for two mechanical rewrites:
  kernel<<<grid, block, smem, stream>>>(args);   ->  emu::launch(kernel, grid, block, smem, stream, args);
  extern __shared__ T name[];                    ->  T* name = (T*)emu::dyn_smem;
  the inline-PTX streaming load of knn.cu        ->  r = *p;
TEST INFRASTRUCTURE (tests/test_rarm_emulated.py)."""
import hashlib
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "retrieval-augmented-diffusion-models_b200", "csrc")
OUT = os.path.join(HERE, "_build")
UNITS = ("rarm.cu", "common.cu", "unet.cu", "kernels.cu", "gemm_simt.cu", "clip.cu", "knn.cu")


def rewrite(src):
    n_launch = len(re.findall(r"<<<", src))
    def launch(m):
        cfg, depth, parts, cur = m.group(2), 0, [], ""
        for ch in cfg:                                   # split the launch configuration at top-level commas; <<<grid, block>>> defaults smem = 0, stream = 0
            depth += ch == "("; depth -= ch == ")"
            if ch == "," and depth == 0:
                parts.append(cur); cur = ""
            else:
                cur += ch
        parts.append(cur)
        parts += ["0", "nullptr"][len(parts) - 2:] if len(parts) < 4 else []
        return f"emu::launch({m.group(1)}, {', '.join(p.strip() for p in parts)}, {m.group(3)});"
    src, k = re.subn(r"([A-Za-z_]\w*(?:<[^<>;()]*>)?)\s*<<<([^;]*?)>>>\s*\(([^;]*?)\);", launch, src)
    assert k == n_launch, f"rewrote {k} of {n_launch} kernel launches"
    src = re.sub(r"extern\s+__shared__\s+(\w+)\s+(\w+)\[\];", r"\1* \2 = (\1*)emu::dyn_smem;", src)
    assert "extern __shared__" not in src
    # the one inline-PTX load of the SIMT units (knn.cu: ld.global.nc.L1::no_allocate, a cache hint) becomes a plain load
    src = re.sub(r'asm volatile\("ld\.global\.nc\.L1::no_allocate\.v4\.u32[^\n]*\n', "r = *p;\n", src)
    src = re.sub(r"\b__noinline__\b", "__attribute__((noinline))", src)       # (a macro of that name would break libstdc++'s own attributes)
    assert "asm volatile" not in src, "inline PTX left in a translation unit that is to be emulated"
    return src


def build(verbose=False):
    os.makedirs(OUT, exist_ok=True)
    inputs = [os.path.join(CSRC, u) for u in UNITS] + [os.path.join(CSRC, "common.cuh"), os.path.join(ROOT, "include", "rdm_b200.h"), os.path.join(HERE, "emu_core.cpp"), os.path.join(HERE, "tc_stubs.cpp"),
                                                       os.path.join(CSRC, "kernels.cuh"), os.path.join(CSRC, "gemm_tc.cuh"), __file__] + [os.path.join(HERE, "include", f) for f in sorted(os.listdir(os.path.join(HERE, "include")))]
    h = hashlib.sha1()
    for p in inputs:
        h.update(open(p, "rb").read())
    lib, stamp = os.path.join(OUT, "librdm_emu.so"), os.path.join(OUT, "stamp")
    if os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == h.hexdigest():
        return lib
    stubs = os.path.join(OUT, "tc_stubs_emu.cpp")
    open(stubs, "w").write(open(os.path.join(HERE, "tc_stubs.cpp")).read().replace('#include "gemm_tc.cuh"', f'#include "{os.path.join(CSRC, "gemm_tc.cuh")}"')
                        .replace('#include "knn_tc.cuh"', f'#include "{os.path.join(CSRC, "knn_tc.cuh")}"'))
    cpps = [os.path.join(HERE, "emu_core.cpp"), stubs]
    for u in UNITS:
        text = rewrite(open(os.path.join(CSRC, u)).read())
        for hdr in ("common.cuh", "kernels.cuh", "gemm_tc.cuh", "knn_tc.cuh"):               # headers of csrc/ by absolute path ("ptx.cuh" resolves to the emulation's)
            text = text.replace(f'#include "{hdr}"', f'#include "{os.path.join(CSRC, hdr)}"')
        text = text.replace('#include "../../include/rdm_b200.h"', f'#include "{os.path.join(ROOT, "include", "rdm_b200.h")}"')
        dst = os.path.join(OUT, u.replace(".cu", "_emu.cpp"))
        open(dst, "w").write(text)
        cpps.append(dst)
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-fno-strict-aliasing", "-Wno-attributes",
           "-I", os.path.join(HERE, "include"), "-o", lib] + cpps
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-6000:], r.stderr[-6000:])
    if r.returncode != 0:
        raise RuntimeError("building the emulated library failed")
    open(stamp, "w").write(h.hexdigest())
    return lib


if __name__ == "__main__":
    print(build(verbose=True))
