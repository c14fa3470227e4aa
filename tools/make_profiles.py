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
for rep, name, lines in (("prof_trace_4k16", "trace_4k16spp_1m", True), ("prof_trace_1080p1", "trace_1080p1spp_1m", True),
                         ("prof_trace_4k16_10m", "trace_4k16spp_10m", True), ("prof_build", "build_kernels_1m", False)):
    path = os.path.join(G, rep + ".ncu-rep")
    if os.path.exists(path):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), path] + (["lines"] if lines else []), capture_output=True, text=True).stdout
        open(os.path.join(P, "%s_%s_ncu.txt" % (tag, name)), "w").write(out)
# 3. hardware counters of the trace kernel for bench.py's roofline block (issue_frac, lanes_active, l1_hit, traffic), tagged with
#    the hash of the kernel source they were captured on: bench.py quotes them only while the source is unchanged
import hashlib
def source_sha():       # must match bench.py:kernel_source_sha (code only: comments and blank lines stripped)
    import re
    h = hashlib.sha256()
    for f in ("csrc/trace.cu", "csrc/bihrt_internal.cuh"):
        src = open(os.path.join(ROOT, "bih-gpu-raytracer_b200", f)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = "\n".join(l.split("//")[0].rstrip() for l in src.splitlines())
        src = "\n".join(l for l in src.splitlines() if l.strip())
        h.update(src.encode())
    return h.hexdigest()[:16]
def counters(rep, rays):
    raw = subprocess.run(["ncu", "-i", os.path.join(G, rep + ".ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines())); h, u, v = rr[0], rr[1], rr[2]
    sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    g = lambda k: float(v[h.index(k)].replace(",", ""))
    rd = g("dram__bytes_read.sum") * sc[u[h.index("dram__bytes_read.sum")]]
    wr = g("dram__bytes_write.sum") * sc[u[h.index("dram__bytes_write.sum")]]
    return {"warp_inst_per_ray": g("smsp__inst_executed.sum") / rays, "lanes_active": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "issue_active_under_ncu": g("smsp__issue_active.avg.per_cycle_active"), "l1_hit": g("l1tex__t_sector_hit_rate.pct") / 100,
            "l2_hit": g("lts__t_sector_hit_rate.pct") / 100, "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
            "rays": rays, "source": "profiles/%s_%s (ncu --set full, one launch)" % (tag, {"prof_trace_4k16": "trace_4k16spp_1m_ncu.txt", "prof_trace_1080p1": "trace_1080p1spp_1m_ncu.txt"}[rep])}
caps = {}
for key, rep, rays in (("bench_4k16spp_1m", "prof_trace_4k16", 3840 * 2160 * 16), ("1080p_1spp_1m", "prof_trace_1080p1", 1920 * 1080)):
    if os.path.exists(os.path.join(G, rep + ".ncu-rep")):
        caps[key] = counters(rep, rays)
json.dump({"kernel_source_sha": source_sha(), "captures": caps}, open(os.path.join(P, "trace_counters.json"), "w"), indent=1)
# 4. SASS of the shipped trace kernel's node step (cuobjdump; the loop between the two votes)
so = os.path.join(ROOT, "bih-gpu-raytracer_b200", "libbihrt.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = sass.split("Function : ")
want = [b for b in blocks if b.startswith("_Z7k_traceILi1ELb0ELi4ELb0EEv9TraceArgs")]
if want:
    lines = [l for l in want[0].splitlines() if "/*" in l and not l.strip().startswith("/* 0x")]
    clean = []
    for l in lines:
        l = l.split("/* 0x")[0].rstrip()
        clean.append(l)
    with open(os.path.join(P, tag + "_k_trace_sass.txt"), "w") as f:
        f.write("# cuobjdump -sass libbihrt.so, function k_trace<1,false,4,false> (camera -> framebuffer, the bench kernel).  %d SASS instructions in all.\n" % len(clean))
        f.write("# The node step is the code between an LDG.E.128.CONSTANT x4 group (node record: planes + child references + two child boxes)\n")
        f.write("# and the next one; the kernel body holds 4 unrolled copies (4 node steps per vote).  No tensor-core or TMA instruction is expected.\n")
        f.write("\n".join(clean) + "\n")
print("profiles written:", sorted(os.listdir(P)))
