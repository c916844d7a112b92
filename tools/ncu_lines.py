#!/usr/bin/env python
"""Join an ncu SASS-level source page with nvdisasm line info and aggregate
executed instructions / stall samples / smem wavefronts per CUDA source line.

usage: ncu_lines.py <report.ncu-rep> <lib.so> <mangled kernel name substring> [top]
"""
import csv, re, subprocess, sys, os, tempfile, collections

rep, so, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
if so.endswith(".hyjit"):
    # a run-time compiled kernel from the cache (hy_jit.hpp): [magic | name length | name | cubin]
    import struct
    blob = open(so, "rb").read()
    nl = struct.unpack("<I", blob[4:8])[0]
    open(os.path.join(tmp, "jit.cubin"), "wb").write(blob[8 + nl:])
elif so.endswith(".cubin"):
    import shutil
    shutil.copy(so, os.path.join(tmp, "jit.cubin"))
else:
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
# (one cubin per translation unit: take the one that holds the kernel)
dis = []
for f in sorted(os.listdir(tmp)):
    if f.endswith(".cubin"):
        out = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if any(ln.startswith("//--------------------- .text.") and kname in ln for ln in out.splitlines()):
            dis = out.splitlines()
            break
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
        cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
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
opagg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
base = None
tot = [0, 0, 0, 0, 0]
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[col["Address"]], 16) if r[col["Address"]].startswith("0x") else int(r[col["Address"]])
    if base is None:
        base = a
    off = a - base
    line, sass = addr2line.get(off, ((None, 0), ""))
    opc = sass.split()[0] if sass and not sass.startswith("@") else (sass.split()[1] if len(sass.split()) > 1 else "?")
    opc = opc.split(".")[0]
    def f(name):
        try:
            return float(r[col[name]] or 0)
        except Exception:
            return 0.0
    v = [f("Instructions Executed"), f("# Samples"), f("L1 Wavefronts Shared"), f("L1 Wavefronts Shared Excessive"), f("Thread Instructions Executed")]
    for i in range(5):
        agg[line][i] += v[i]
        opagg[opc][i] += v[i]
        tot[i] += v[i]
srcdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "heyoka.py_b200", "csrc")
srcs = {}
def src_line(fl):
    f, l = fl
    if f is None:
        return "?"
    if f not in srcs:
        try:
            srcs[f] = open(os.path.join(srcdir, f)).read().splitlines()
        except Exception:
            srcs[f] = []
    t = srcs[f]
    return "%s:%d  %s" % (f, l, t[l - 1].strip()[:80] if 0 < l <= len(t) else "")
print("total inst %.3g samples %d smem wavefronts %.3g excessive %.3g" % (tot[0], tot[1], tot[2], tot[3]))
print("%6s %7s %7s %7s %7s  %s" % ("line", "inst%", "samp%", "wf%", "exc%", "source"))
for line, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    s = src_line(line)
    print("%6s %7.2f %7.2f %7.2f %7.2f  %s" % ("", 100 * v[0] / tot[0], 100 * v[1] / max(tot[1], 1), 100 * v[2] / max(tot[2], 1), 100 * v[3] / max(tot[3], 1), s))
print("per opcode: inst% samp% active-threads/inst")
for opc, v in sorted(opagg.items(), key=lambda kv: -kv[1][0])[:25]:
    print("%12s %7.2f %7.2f %6.1f" % (opc, 100 * v[0] / tot[0], 100 * v[1] / max(tot[1], 1), v[4] / max(v[0], 1)))
