"""Turn gpurun_out/{launches,prof}_<tag> into the tracked summaries under profiles/."""
import collections, csv, json, re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = ROOT / "profiles"
out.mkdir(exist_ok=True)
lst = ROOT / "gpurun_out" / f"launches_{tag}.csv"
if lst.exists():
    lines = [l for l in lst.read_text().splitlines() if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        m = re.search(r"(k_[a-z_]+)", row["Kernel Name"])
        k = m.group(1) if m else row["Kernel Name"][:40]
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        agg[k][0] += 1
        agg[k][1] += v
        tot += v
    with open(out / f"{tag}_launch_list_summary.txt", "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 1 --warmup 1 --no-cpu-baseline\n")
        f.write(f"# serialised, cold-cache launch times: compare SHARES, not absolutes.  total {tot / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k:22s} launches={n:5d} total_ms={t / 1e3:10.3f} avg_us={t / n:9.1f} share={100 * t / tot:5.1f}%\n")
    print((out / f"{tag}_launch_list_summary.txt").read_text())
rep = ROOT / "gpurun_out" / f"prof_{tag}.ncu-rep"
if rep.exists():
    txt = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(rep)], capture_output=True, text=True).stdout
    (out / f"{tag}_ncu_full_summary.txt").write_text(
        "# ncu --set full --clock-control none --import-source on, kernels of one bench step (C2, 1080p, bench.py default pool)\n" + txt)
    print(txt)
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    idx = {h: i for i, h in enumerate(rows[0])}
    units = rows[1]
    tr = []
    kkey = "k_trace_fast" if any("k_trace_fast" in d[idx["Kernel Name"]] for d in rows[2:]) else "k_extend_fast"
    for d in rows[2:]:
        if kkey in d[idx["Kernel Name"]]:
            def b(name):
                v = float(d[idx[name]].replace(",", ""))
                u = units[idx[name]]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            tr.append((b("dram__bytes_read.sum") + b("dram__bytes_write.sum"), float(d[idx["lts__t_sectors.sum"]].replace(",", "")) * 32.0,
                       float(d[idx["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(units[idx["gpu__time_duration.sum"]], 1e-6)))
    if tr:
        n = len(tr)
        (out / "traffic.json").write_text(json.dumps({
            kkey + "_dram_bytes_per_launch": sum(t[0] for t in tr) / n,
            kkey + "_l2_bytes_per_launch": sum(t[1] for t in tr) / n,
            kkey + "_ms_under_ncu": sum(t[2] for t in tr) / n,
            "launches_sampled": n, "what": "full-pool (32 Mi slots) launches of one 64-spp C2 step, ncu --set full",
            "source": f"profiles/{tag}_ncu_full_summary.txt"}))
        print("traffic", tr)
