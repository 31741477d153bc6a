"""Per-kernel SASS opcode histogram of the product library (no GPU needed).
usage: sass_hist.py [library or object] [kernel regex] > profiles/<tag>_sass_histogram.txt
Lists, per kernel, the instruction count and the opcodes that show what the code was written for on sm_100a:
packed FP32 (FADD2 / FMUL2 / FFMA2), 3-input min/max (FMNMX3, VIMNMX3), TMA bulk copies (UBLKCP) with their
mbarriers (SYNCS), 256-bit global loads (LDG.E.*256), warp votes / shuffles / reductions (VOTE, SHFL, REDUX),
byte permutes (PRMT) of the compressed-node decode -- and that there is no tensor-core instruction (nothing on this
path is a contraction)."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
target = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "tryraytrace_b200" / "lib" / "libtrt_b200.so")
kre = re.compile(sys.argv[2]) if len(sys.argv) > 2 else re.compile(r"k_trace_fastILi768ELb0ELb[01]ELb0|k_shadeILb0ELb1ELi128ELi8|k_refill|k_finish_pathsILb0|k_compress_nodes|k_extend_fastILi768ELb0ELb0|k_shadow_fastILi768ELb0ELb0|k_compact_move|k_instance_objects")
sass = subprocess.run(["cuobjdump", "-sass", target], capture_output=True, text=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
KEY = ["FADD2", "FMUL2", "FFMA2", "FMNMX3", "VIMNMX3", "UBLKCP", "SYNCS", "LDG.E.ENL2.256", "LDG.E.128", "STG.E.128", "PRMT",
       "VOTE", "SHFL", "REDUX", "LDS", "STS", "CCTL", "MUFU", "HMMA", "UTCMMA", "LDTM"]
print(f"# cuobjdump -sass {Path(target).name} | tools/sass_hist.py   (sm_100a; opcode counts are static instructions)")
tot = collections.Counter()
for name, c in kernels.items():
    if not kre.search(name):
        continue
    demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    mm = re.search(r"(k_\w+(?:<[^>]*>)?)", demangled)
    short = mm.group(1) if mm else demangled[:60]
    n = sum(c.values())
    parts = []
    for k in KEY:
        v = sum(cnt for op, cnt in c.items() if op == k or op.startswith(k + "."))
        if v:
            parts.append(f"{k}={v}")
            tot[k] += v
    top = ", ".join(f"{op} {cnt}" for op, cnt in c.most_common(8))
    print(f"{short}\n    instructions={n}  " + "  ".join(parts) + f"\n    most frequent: {top}")
print("# totals over the listed kernels: " + "  ".join(f"{k}={v}" for k, v in tot.items()))
allc = collections.Counter()
for c in kernels.values():
    allc.update(c)
tc = sum(v for op, v in allc.items() if re.match(r"(HMMA|IMMA|DMMA|BMMA|UTC\w*MMA|LDTM|STTM|UTCBAR)", op))
print(f"# whole library: {sum(allc.values())} instructions in {len(kernels)} kernels; tensor-core / TMEM instructions: {tc}")
