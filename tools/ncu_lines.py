"""Per-source-line rollup of an ncu capture (no GPU needed).

ncu's `--page source --csv` export lists SASS instructions with their executed counts and stall
samples but no line numbers; `nvdisasm -g` of the same cubin lists the same instructions with
`//## File "...", line N [inlined at ...]` markers.  Both are in address order, so the two are joined
by instruction index, and executed warp instructions / thread instructions / stall samples are
summed per source line (innermost inlined location).

usage: ncu_lines.py <report.ncu-rep> <kernel regex> <object or cubin with -lineinfo> [top N] [launch index] [inner|outer|chain]
  inner (default): group by the innermost location; outer: by the line of the kernel body the code was inlined
  into; chain: by the whole inline chain.
"""
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

rep, kre, obj = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
which = int(sys.argv[5]) if len(sys.argv) > 5 else 0
by = sys.argv[6] if len(sys.argv) > 6 else "inner"

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
# the export holds one block per launch: "Kernel Name",<name> then a header row then instruction rows
blocks, cur = [], None
for row in csv.reader(raw.splitlines()):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None and row and row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] is not None and row and row[0].startswith("0x"):
        cur["rows"].append(row)
if not blocks:
    sys.exit("no kernel matched")
blk = blocks[which]
hdr = {h: i for i, h in enumerate(blk["hdr"])}
name = blk["name"]

# mangled name of that kernel inside the cubin: match on template arguments is fragile, so match by
# instruction count and opcode sequence against every function of the disassembly
with tempfile.TemporaryDirectory() as td:
    p = Path(obj)
    if p.suffix != ".cubin":
        subprocess.run(["cuobjdump", "-xelf", "all", str(p.resolve())], cwd=td, check=True, capture_output=True)
        p = next(Path(td).glob("*.cubin"))
    dis = subprocess.run(["nvdisasm", "-gi", "-c", str(p)], capture_output=True, text=True).stdout

funcs, f = {}, None
loc, fresh = None, True
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        f = funcs.setdefault(m.group(1), [])
        loc, fresh = None, True
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:  # consecutive marker lines = one inline chain, innermost first
        here = f"{Path(m.group(1)).name}:{m.group(2)}"
        chain = [here] if fresh else list(loc[2]) + [here]
        fresh = False
        key = chain[0] if by == "inner" else (chain[-1] if by == "outer" else " < ".join(chain))
        loc = (key, 0, tuple(chain))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and f is not None:
        fresh = True
        f.append((m.group(2).split()[0 if not m.group(2).startswith("@") else 1], loc))

ops = [re.sub(r"^@!?U?P\d+\s+", "", r[hdr["Source"]].strip()).split()[0] for r in blk["rows"]]
cands = [k for k, v in funcs.items() if len(v) == len(ops) and all(a[0] == b for a, b in zip(v, ops))]
if not cands:
    sys.exit(f"no function of {obj} matches the {len(ops)} instructions of {name[:80]} (rebuilt since the capture?)")
locs = [l for _, l in funcs[cands[0]]]

inner_file = {}
tot = defaultdict(lambda: [0, 0, 0, 0])  # warp inst, thread inst, stall samples, static inst
for r, l in zip(blk["rows"], locs):
    a = tot[l[0] if l else '?']
    if l: inner_file[l[0]] = l[2][0].split(':')[0]
    a[0] += int(r[hdr["Instructions Executed"]] or 0)
    a[1] += int(r[hdr["Thread Instructions Executed"]] or 0)
    a[2] += int(r[hdr["# Samples"]] or 0)
    a[3] += 1
W = sum(a[0] for a in tot.values()) or 1
S = sum(a[2] for a in tot.values()) or 1
print(f"# {name[:110]}")
print(f"# {len(ops)} SASS instructions, {W:.4g} warp instructions executed, {S} stall samples")
print(f"# {'file:line':28s} {'warp inst':>11s} {'share':>6s} {'lanes':>6s} {'samples':>8s} {'share':>6s} static")
for l, a in sorted(tot.items(), key=lambda kv: -kv[1][2])[:top]:
    where = l
    print(f"  {where:{28 if by != 'chain' else 90}s} {a[0]:11d} {100 * a[0] / W:5.1f}% {a[1] / max(a[0], 1):6.1f} {a[2]:8d} {100 * a[2] / S:5.1f}% {a[3]:5d}")
# per file
pf = defaultdict(lambda: [0, 0])
for l, a in tot.items():
    k = inner_file.get(l, '?') if by == 'inner' else l.split(':')[0]
    pf[k][0] += a[0]
    pf[k][1] += a[2]
print("# per file: " + "  ".join(f"{k} inst {100 * v[0] / W:.1f}% samples {100 * v[1] / S:.1f}%" for k, v in sorted(pf.items(), key=lambda kv: -kv[1][1])))
