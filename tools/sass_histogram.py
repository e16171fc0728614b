"""SASS instruction histogram of the default kernel instantiations (CPU only:
cuobjdump on the objects of the in-tree build).  Evidence for how the hot
kernels are built: LDG.E.128 streams, F2F / I2F conversions, DFMA, the PDL
instructions (ACQBULK / PREEXIT), the cluster / DSMEM instructions of the TRSV
kernel; no tensor-core instruction by design (north star).

    python tools/sass_histogram.py > profiles/r02_sass.md
"""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
BUILD = ROOT / "build" / "accblas"

# (object, regex on the demangled kernel name, label)
KERNELS = []
for st, stn in (("double", "fp64"), ("float", "fp32"), ("__half", "fp16")):
    for ar, arn in (("double", "fp64"), ("float", "fp32")):
        fast = st == "__half" and ar == "double"
        KERNELS.append(("gemv.o",
                        rf"gemv_stream_kernel<{st}, {ar}, \(int\)4, \(int\)2, \(int\)1, \(int\)8, \(int\)3, "
                        + (r"\(int\)1, \(int\)2, \(int\)8>" if fast else r"\(int\)0, \(int\)2, \(int\)16>"),
                        f"GEMV Acc<{arn},{stn}>"))
for st, stn in (("double", "fp64"), ("float", "fp32"), ("__half", "fp16")):
    for ar, arn in (("double", "fp64"), ("float", "fp32")):
        block = 1024 if (st == "double" and ar == "double") else 256
        KERNELS.append(("dot.o",
                        rf"dot_stream_kernel<{st}, {ar}, \(int\){block}, \(int\)4, \(int\)16, \(bool\)0>",
                        f"DOT Acc<{arn},{stn}>"))
for obj, ar in (("trsv_cluster_f64.o", "double"), ("trsv_cluster_f32.o", "float")):
    for st, stn, vw in (("double", "fp64", 16), ("float", "fp32", 16), ("__half", "fp16", 8)):
        KERNELS.append((obj,
                        rf"trsv_cluster_kernel<{st}, {ar}, \(bool\)0, \(bool\)1, \(int\){vw}, \(bool\)0>",
                        f"TRSV cluster lower/unit Acc<{'fp64' if ar == 'double' else 'fp32'},{stn}>"))
KERNELS.append(("trsv.o", r"trsv_kernel<float, double, \(bool\)0, \(bool\)1, \(int\)16, \(bool\)0>",
                "TRSV single-CTA lower/unit Acc<fp64,fp32>"))
KERNELS.append(("convert_fill.o", r"fill_linear_kernel<float>", "fill_uniform fp32 (linear)"))
KERNELS.append(("convert_fill.o", r"convert_kernel<float, double, \(bool\)1>", "convert fp64 -> fp32"))

INTEREST = ["LDG", "STG", "LDS", "STS", "LDGSTS", "UBLKCP", "F2F", "I2F", "F2FP", "HADD2", "HFMA2", "DFMA",
            "DADD", "DMUL", "FFMA", "FMUL", "IMAD", "LOP3", "SHFL", "BAR", "ACQBULK", "PREEXIT", "SYNCS", "MAPA",
            "UCGABAR", "STAS", "ATOM", "RED", "STL", "LDL", "HMMA", "UTC", "LDTM"]


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", str(BUILD / obj)], capture_output=True, text=True).stdout
    demangled = subprocess.run(["cu++filt"], input=out, capture_output=True, text=True).stdout
    cur, res = None, {}
    for line in demangled.splitlines():
        m = re.match(r"\s*Function : (.*)", line)
        if m:
            cur = m.group(1)
            res[cur] = Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            res[cur][m.group(1)] += 1
    return res


def main():
    cache = {}
    print("# SASS instruction histograms of the default instantiations (sm_100a)\n")
    print("`python tools/sass_histogram.py` on the objects of the in-tree build (`cuobjdump -sass build/accblas/*.o`). "
          "Counts are static instructions per kernel. No tensor-core instruction (`UTC*MMA`, `HMMA`, `LDTM`) appears: "
          "the path is HBM-bound matrix-vector work by design. `ACQBULK`/`PREEXIT` = programmatic dependent launch; "
          "`UCGABAR_*`, `MAPA`, `STAS` (st.async) and `SYNCS` (mbarrier) = thread-block clusters and distributed shared "
          "memory in the TRSV kernel.\n")
    for obj, pattern, label in KERNELS:
        if obj not in cache:
            cache[obj] = functions(obj)
        hits = [(name, c) for name, c in cache[obj].items() if re.search(pattern, name)]
        if not hits:
            print(f"## {label}\n\n(not found: `{pattern}`)\n")
            continue
        name, c = hits[0]
        total = sum(c.values())
        groups = Counter()
        for op, cnt in c.items():
            for key in INTEREST:
                if op.startswith(key):
                    groups[key] += cnt
                    break
        wide = sum(cnt for op, cnt in c.items() if op.startswith("LDG") and ".128" in op)
        print(f"## {label}\n\n`{name[:200]}`\n")
        print(f"{total} instructions; 128-bit global loads: {wide}\n")
        print("| " + " | ".join(k for k in INTEREST if groups[k]) + " |")
        print("|" + "---|" * sum(1 for k in INTEREST if groups[k]))
        print("| " + " | ".join(str(groups[k]) for k in INTEREST if groups[k]) + " |\n")
        top = ", ".join(f"{op} {cnt}" for op, cnt in c.most_common(12))
        print(f"most frequent: {top}\n")


if __name__ == "__main__":
    main()
