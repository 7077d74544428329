"""Static SASS / resource table of the kernels in liblncde_b200.so (no GPU needed): registers, shared memory, stack,
instruction count and mix (FP32 arithmetic, shared / global memory, block barriers, shuffles) and the Blackwell-specific
mnemonics (UBLKCP = 1-D TMA bulk copy, UTMALDG = tensor-map TMA load, UTCHMMA = tcgen05.mma, HMMA = warp-level MMA,
LDGSTS = cp.async, SYNCS = mbarrier, LDTM = tcgen05.ld, STAS = st.async into a peer CTA's shared memory, UCGABAR = cluster
barrier, FFMA2 = packed fp32 pair).  Counts are STATIC (whole kernel, every instantiated path once), so they compare
variants of the same kernel - e.g. block barriers of the default and the fused-phase Log-NCDE kernels - not run times.

    python scripts/sass_stats.py [regex] > profiles/r02_static_sass.md
"""
import os
import re
import subprocess
import sys
from collections import Counter

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(R, "log-neural-cdes_b200", "liblncde_b200.so")
CLASSES = (("fp32", r"^(FFMA|FMUL|FADD|FMNMX|MUFU|FSEL|FSET|FSETP)"), ("lds", r"^LDS"), ("sts", r"^STS"),
           ("ldg", r"^(LDG|LD\.)"), ("stg", r"^(STG|ST\.)"), ("bar", r"^BAR"), ("shfl", r"^SHFL"))
SPECIAL = ("UBLKCP", "UTMALDG", "UTMASTG", "UTCHMMA", "UTCBAR", "LDTM", "HMMA", "LDGSTS", "SYNCS", "STAS", "UCGABAR", "FFMA2")
DEFAULT_PICK = (r"logsig_levy(_tma)?_kernel<7, ?(true|1)|logsig_d3_kernel<7>|logncde_(fwd_fast|bwd_saved)_kernel<7|"
                r"outer_gemm_kernel|flow_build|dlie_t|scan_(fwd|bwd|lam)_(warp_kernel<4|tma|cluster_kernel<4)|pscan_\w+<4|dlncde_\w+<7>|"
                r"nrde_(fwd|bwd)_kernel|umma_|cutlass|adam_dev|ffma_probe")


def main(pattern):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage, cur = {}, None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and cur:
            usage[cur] = tuple(int(v) for v in m.groups())
    funcs, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            funcs[cur]["n"] += 1
            for name, rx in CLASSES:
                if re.match(rx, op):
                    funcs[cur][name] += 1
            for sp in SPECIAL:
                if op.startswith(sp):
                    funcs[cur][sp] += 1
    names = list(funcs)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    rows = []
    for n, d in zip(names, dem):
        short = re.sub(r"\(.*", "", d.replace("lncde::fast::", "fast::").replace("lncde::", "").replace("void ", ""))
        if short.startswith("_ZN7cutlass") or short.startswith("cutlass"):  # c++filt gives up on the 2 KB template name
            kind = "FastF32" if "FastF32" in n else "?"
            short = f"cutlass::device_kernel<GemmUniversal<MainloopSm100TmaUmmaWarpSpecialized{kind}, ...>> #{len(rows)}"
        if len(short) > 90:
            short = short[:87] + "..."
        if re.search(pattern, d):
            rows.append((short, usage.get(n, (0, 0, 0)), funcs[n]))
    rows.sort()
    print("# Static SASS / resource table (`python scripts/sass_stats.py`, sm_100a build of this tree)\n")
    print("Static counts over the whole kernel - they compare variants, they are not run times.  `special` = Blackwell /\n"
          "async mnemonics found (UBLKCP: 1-D TMA bulk copy, UTMALDG: tensor-map TMA, UTCHMMA: tcgen05.mma, HMMA: warp MMA,\n"
          "LDGSTS: cp.async, SYNCS: mbarrier).\n")
    print("| kernel | regs | stack B | static smem B | instr | fp32 | LDS | STS | LDG | STG | BAR | SHFL | special |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
    for short, (reg, stack, smem), c in rows:
        sp = ", ".join(f"{k} {c[k]}" for k in SPECIAL if c[k])
        print(f"| `{short}` | {reg} | {stack} | {smem} | {c['n']} | {c['fp32']} | {c['lds']} | {c['sts']} | {c['ldg']} | "
              f"{c['stg']} | {c['bar']} | {c['shfl']} | {sp} |")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else DEFAULT_PICK)
