"""Rank source lines of an ncu report by executed warp instructions (development aid).
usage: python tools/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = ""; hdr = None; lines = []; tot = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ci = hdr.index("Instructions Executed"); ti = hdr.index("Thread Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr and r[0].strip().isdigit():
        try: ie = int(r[ci]); te = int(r[ti]); sm = int(r[si])
        except Exception: continue
        lines.append((fname, int(r[0]), r[1].strip()[:100], ie, te, sm)); tot += ie
print("total warp instructions", tot, " thread instr", sum(l[4] for l in lines), " avg active %.1f" % (sum(l[4] for l in lines) / max(tot, 1)))
tsm = sum(l[5] for l in lines)
for f, ln, src, ie, te, sm in sorted(lines, key=lambda x: -x[3])[:top]:
    print("%-14s %4d %5.1f%% act %4.1f smp %4.1f%% | %s" % (f[:14], ln, 100 * ie / tot, te / max(ie, 1), 100 * sm / max(tsm, 1), src))
