#!/usr/bin/env python
"""Per-CUDA-source-line view of one kernel: joins the SASS execution counts of an .ncu-rep
(source page) with the line table of the built library (nvdisasm -g), by instruction offset.
usage: tools/ncu_lines.py rep kernel-regex launch-idx cubin-or-so function-substring [--ops]
Prints, per source line, warp-level instructions executed / stall samples, and the share of
the fp64 pipe (DADD/DFMA/DMUL/DSETP)."""
import csv, glob, io, os, re, subprocess, sys, tempfile

rep, rx, idx, lib, fsub = sys.argv[1:6]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", "::regex:%s:%s" % (rx, idx)],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
if len(rows) < 3:   # some ncu versions reject the kernel-id filter on imported reports: take the first kernel
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ins = []
for r in rows[2:]:
    if len(r) <= ismp or not r[ia].startswith("0x"):
        if ins:
            break
        continue
    ins.append((int(r[ia], 16), r[isrc].strip(), int(r[iex]), int(r[ismp])))
base = ins[0][0]

if lib.endswith(".so"):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cubins = glob.glob(os.path.join(tmp, "*.cubin"))
else:
    cubins = [lib]
line_of = {}
src_file = None
for cb in cubins:
    out = subprocess.run(["nvdisasm", "-g", "-c", cb], capture_output=True, text=True).stdout
    infn, cur = False, None
    for l in out.splitlines():
        if l.startswith(".text."):
            infn = (fsub in l) and l.rstrip().endswith(":")
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
        if m:
            line_of[int(m.group(1), 16)] = cur
    if line_of:
        break

FP64 = ("DADD", "DFMA", "DMUL", "DSETP")
agg = {}
tot_e = sum(i[2] for i in ins)
tot_s = sum(i[3] for i in ins)
for a, s, e, m in ins:
    ln = line_of.get(a - base)
    op = s.split()[0] if not s.startswith("@") else s.split()[1]
    op = op.split(".")[0]
    d = agg.setdefault(ln, {"e": 0, "s": 0, "f": 0, "ops": {}})
    d["e"] += e
    d["s"] += m
    if op in FP64:
        d["f"] += e
    d["ops"][op] = d["ops"].get(op, 0) + e
srcs = {}
def text(ln):
    if ln is None:
        return "?"
    f, n = ln
    if f not in srcs:
        try:
            srcs[f] = open(f).read().splitlines()
        except Exception:
            srcs[f] = []
    t = srcs[f][n - 1].strip() if n - 1 < len(srcs[f]) else ""
    return "%s:%d  %s" % (os.path.basename(f), n, t[:90])
print("total warp-instr %d, samples %d, fp64 %.1f%%" % (tot_e, tot_s, 100.0 * sum(d["f"] for d in agg.values()) / tot_e))
for ln, d in sorted(agg.items(), key=lambda kv: (kv[0] is None, kv[0])):
    if d["e"] * 1000 < tot_e and d["s"] * 1000 < tot_s:
        continue
    extra = ""
    if "--ops" in sys.argv:
        extra = "  " + " ".join("%s:%.2f" % (k, 100.0 * v / tot_e) for k, v in sorted(d["ops"].items(), key=lambda kv: -kv[1])[:5])
    print("%5.2f%% exec %5.2f%% smp %5.2f%% f64 | %s%s" % (100.0 * d["e"] / tot_e, 100.0 * d["s"] / max(tot_s, 1),
                                                        100.0 * d["f"] / tot_e, text(ln), extra))
