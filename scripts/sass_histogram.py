"""Per-kernel SASS opcode histogram of libnerfloam_b200.so (cuobjdump -sass): the evidence that the hot kernels are tcgen05 / TMEM /
bulk-copy / multimem code.  Writes profiles/r02_sass_histogram.md.  Runs without a GPU."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "nerf-loam_b200", "libnerfloam_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kern, hist = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = m.group(1); hist[kern] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if m and kern:
        hist[kern][m.group(1).split(".")[0] if not m.group(1).startswith(("UTC", "LDGMC", "STGMC", "RED", "ATOM", "UBLKCP", "LDTM", "STTM", "SYNCS")) else m.group(1)] += 1
KEY = ("UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "UTMA", "SYNCS", "LDGMC", "REDG", "RED", "ATOMG", "ATOMS", "HMMA", "FFMA", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "ELECT")
lines = ["# SASS opcode histogram per kernel (round 2 build)", "",
         "`python scripts/sass_histogram.py` = `cuobjdump -sass nerf-loam_b200/libnerfloam_b200.so`, instruction mnemonics counted per `Function`.",
         "Columns: selected opcode families (prefix match); `total` = all instructions of the kernel.  UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM),",
         "UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier ops, LDGMC = multimem.ld_reduce (NVLS; multimem.st compiles to a plain STG.E.128.STRONG.SYS on the multicast address), REDG = red.global (fp32x4 atomics).", "",
         "| kernel | total | " + " | ".join(KEY) + " |", "|---|---:|" + "---:|" * len(KEY)]
tot = collections.Counter()
for k, h in hist.items():
    name = demangle(k)
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*", "", name)[:70]
    row = [sum(v for op, v in h.items() if op.startswith(key)) for key in KEY]
    for key, v in zip(KEY, row): tot[key] += v
    lines.append(f"| `{name}` | {sum(h.values())} | " + " | ".join(str(v) if v else "" for v in row) + " |")
lines.append("| **all kernels** | | " + " | ".join(str(tot[k]) for k in KEY) + " |")
open(os.path.join(ROOT, "profiles", "r02_sass_histogram.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[-12:]))
