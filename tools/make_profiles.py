"""Turns the ncu captures in gpurun_out/ into the tracked summaries under profiles/ (round tag as argv[1])."""
import csv, collections, json, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
# 1. launch list of the bench command
rows = list(csv.reader(open(os.path.join(G, "launches.csv"))))
hdr = None; agg = collections.OrderedDict(); order = []
for r in rows:
    if hdr is None:
        if "Kernel Name" in r: hdr = r
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") == "gpu__time_duration.sum":
        v = float(d["Metric Value"].replace(",", "")); u = d["Metric Unit"]
        v = v / 1e3 if u.startswith("n") else (v * 1e3 if u.startswith("m") else v)      # -> us
        order.append((d["Kernel Name"][:70], v))
        a = agg.setdefault(d["Kernel Name"][:70], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in order)
with open(os.path.join(P, tag + "_bench_launches.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 3 --no-extras\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes.  total %.1f us over %d launches\n" % (tot, len(order)))
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write("%-72s n=%3d total %10.1f us  avg %10.1f us  share %5.1f%%\n" % (k, c, t, t / c, 100 * t / tot))
    f.write("\n# in launch order\n")
    for k, v in order:
        f.write("%-72s %10.1f us\n" % (k, v))
# 2. full captures
for rep, name, lines in (("prof_trace_4k16", "trace_4k16spp_1m", True), ("prof_trace_1080p1", "trace_1080p1spp_1m", True), ("prof_build", "build_kernels_1m", False)):
    path = os.path.join(G, rep + ".ncu-rep")
    if os.path.exists(path):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), path] + (["lines"] if lines else []), capture_output=True, text=True).stdout
        open(os.path.join(P, "%s_%s_ncu.txt" % (tag, name)), "w").write(out)
# 3. DRAM traffic of the dominant kernel for bench.py's roofline.traffic
raw = subprocess.run(["ncu", "-i", os.path.join(G, "prof_trace_4k16.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); h, u, v = rr[0], rr[1], rr[2]
sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd = float(v[h.index("dram__bytes_read.sum")]) * sc[u[h.index("dram__bytes_read.sum")]]
wr = float(v[h.index("dram__bytes_write.sum")]) * sc[u[h.index("dram__bytes_write.sum")]]
json.dump({"kernel": "k_trace<1,0> (bench workload: 1 M triangles, 3840x2160 x 16 spp)", "dram_bytes_per_launch": rd + wr,
           "dram_bytes_read": rd, "dram_bytes_write": wr, "source": "%s_trace_4k16spp_1m_ncu.txt (ncu --set full, one launch)" % tag},
          open(os.path.join(P, "trace_traffic.json"), "w"), indent=1)
print("profiles written:", sorted(os.listdir(P)))
