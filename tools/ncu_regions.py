"""Per-SASS-line listing of one kernel from an .ncu-rep source page, with hot regions.
usage: ncu_regions.py report.ncu-rep kernel-regex [min_kinst]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
mink = float(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}",
                      "--launch-skip", "0", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data, seen = [], set()
for r in rows[2:]:
    if len(r) < len(hdr) - 2 or not r[0].startswith("0x"):
        continue
    if r[0] in seen:
        break
    seen.add(r[0]); data.append(r)
tot_s = sum(int(r[ix["# Samples"]]) for r in data)
tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data)
print("sass lines", len(data), "samples", tot_s, "warp inst", tot_i)
for k, r in enumerate(data):
    s = int(r[ix["# Samples"]]); ie = int(r[ix["Instructions Executed"]])
    if ie / 1e3 < mink:
        continue
    print(f"{k:4d} {100*s/max(tot_s,1):5.2f}% {ie/1e3:8.0f}k thr={float(r[ix['Avg. Threads Executed']]):4.1f} "
          f"lsb={r[ix['stall_long_sb']]:>5} ssb={r[ix['stall_short_sb']]:>5} {r[ix['Source']].strip()[:80]}")
