#!/usr/bin/env python
"""Join an ncu SASS-level source page with nvdisasm line info and aggregate
executed instructions / stall samples / smem wavefronts per CUDA source line.

usage: ncu_lines.py <report.ncu-rep> <lib.so> <mangled kernel name substring> [top]
"""
import csv, re, subprocess, sys, os, tempfile, collections

rep, so, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
# locate function text section
addr2line = {}
cur_line = None
infn = False
for ln in dis:
    if ln.startswith("//--------------------- .text."):
        infn = kname in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = int(m.group(2)) if m.group(1).endswith("hy_kernels.cuh") else -1
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m:
        addr2line[int(m.group(1), 16)] = (cur_line, m.group(2))
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
base = None
tot = [0, 0, 0, 0, 0]
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[col["Address"]], 16) if r[col["Address"]].startswith("0x") else int(r[col["Address"]])
    if base is None:
        base = a
    off = a - base
    line = addr2line.get(off, (None, ""))[0]
    def f(name):
        try:
            return float(r[col[name]] or 0)
        except Exception:
            return 0.0
    v = [f("Instructions Executed"), f("# Samples"), f("L1 Wavefronts Shared"), f("L1 Wavefronts Shared Excessive"), f("Thread Instructions Executed")]
    for i in range(5):
        agg[line][i] += v[i]
        tot[i] += v[i]
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "heyoka.py_b200", "csrc", "hy_kernels.cuh")).read().splitlines()
print("total inst %.3g samples %d smem wavefronts %.3g excessive %.3g" % (tot[0], tot[1], tot[2], tot[3]))
print("%6s %7s %7s %7s %7s  %s" % ("line", "inst%", "samp%", "wf%", "exc%", "source"))
for line, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    s = src[line - 1].strip()[:90] if line and 0 < line <= len(src) else str(line)
    print("%6s %7.2f %7.2f %7.2f %7.2f  %s" % (line, 100 * v[0] / tot[0], 100 * v[1] / max(tot[1], 1), 100 * v[2] / max(tot[2], 1), 100 * v[3] / max(tot[3], 1), s))
