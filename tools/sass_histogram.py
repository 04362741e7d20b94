"""Opcode histogram of the Blackwell-specific SASS in librdm_b200.so, per kernel: the tcgen05 / TMEM / TMA claim must not
depend on a reader rebuilding the library.   python tools/sass_histogram.py > profiles/sass_histogram_r2.md
Mnemonics (B200_PROFILING.md): UTCHMMA / UTCQMMA = tcgen05.mma (f16 / block-scaled kinds), LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG / UTMASTG = TMA tensor
loads / stores (cp.async.bulk.tensor), UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier ops, UTCBAR = tcgen05.commit, HMMA = mma.sync (warp MMA),
REDUX / SHFL = warp reductions / shuffles, ACQBULK / UTCATOMSWS = TMEM allocator."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "retrieval-augmented-diffusion-models_b200", "rdm_b200", "librdm_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA", "DFMA", "DADD", "DMUL", "FFMA", "MUFU", "SHFL", "ATOM", "ATOMS", "RED", "LDGSTS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per, total, cur = collections.OrderedDict(), collections.Counter(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(.*$", "", name)
            cur = per.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_all"] += 1
            if op in WATCH:
                cur[op] += 1; total[op] += 1
    print("# SASS opcode histogram of `librdm_b200.so` (sm_100a), round 2\n")
    print("`python tools/sass_histogram.py` = `cuobjdump -sass` of the in-tree library, instructions counted per kernel (static counts, not executed counts).\n")
    print("Whole library: " + ", ".join(f"{total[o]} `{o}`" for o in WATCH if total[o]) + "\n")
    cols = [o for o in WATCH if total[o]]
    print("| kernel | SASS instr | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    agg = collections.OrderedDict()
    for name, c in per.items():
        base = re.sub(r"<.*$", "", name)
        a = agg.setdefault(base, [0, collections.Counter()])
        a[0] += 1; a[1].update(c)
    for base, (n, c) in sorted(agg.items(), key=lambda kv: -(kv[1][1]["UTCHMMA"] * 1000 + kv[1][1]["UTMALDG"] * 100 + kv[1][1]["UBLKCP"] * 10 + kv[1][1]["HMMA"])):
        label = f"`{base}`" + (f" ({n} instantiations, summed)" if n > 1 else "")
        print(f"| {label} | {c['_all']} | " + " | ".join(str(c[o]) if c[o] else "" for o in cols) + " |")


if __name__ == "__main__":
    main()
