"""Group the SASS lines of one kernel (ncu source page) into blocks of equal execution count.
usage: ncu_blocks.py report.ncu-rep kernel-regex [min_pct]"""
import re, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run([sys.executable, __file__.replace("ncu_blocks", "ncu_regions"), rep, rx, "0"], capture_output=True, text=True).stdout
rows = []
for l in out.splitlines()[1:]:
    m = re.match(r'\s*(\d+)\s+([\d.]+)%\s+(\d+)k thr=\s*([\d.]+) lsb=\s*(\d+) ssb=\s*(\d+) (.*)', l)
    if m:
        rows.append((int(m[1]), float(m[2]), int(m[3]), float(m[4]), m[7]))
tot = sum(r[2] for r in rows)
blk, cur = [], None
for r in rows:
    if cur and abs(r[2] - cur['k']) <= 0.15 * max(cur['k'], 50):
        cur['n'] += 1; cur['inst'] += r[2]; cur['samp'] += r[1]; cur['thr'] += r[3] * r[2]; cur['end'] = r[0]
    else:
        if cur: blk.append(cur)
        cur = {'start': r[0], 'end': r[0], 'k': r[2], 'n': 1, 'inst': r[2], 'samp': r[1], 'thr': r[3] * r[2], 'first': r[4]}
blk.append(cur)
print("total warp inst (k):", tot)
for b in blk:
    if b['inst'] > minp / 100 * tot:
        print(f"{b['start']:5d}-{b['end']:5d} n={b['n']:4d} exec/line={b['k']:7d}k inst={100*b['inst']/tot:5.1f}% samp={b['samp']:5.1f}% thr={b['thr']/max(b['inst'],1):4.1f}  {b['first'][:50]}")
