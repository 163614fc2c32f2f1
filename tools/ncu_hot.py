#!/usr/bin/env python
"""Per-instruction view of one kernel from an .ncu-rep source page: executed counts and stall
samples in SASS order, bucketed in blocks of N instructions.
usage: tools/ncu_hot.py rep kernel-regex [launch-idx] [--sass]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
idx = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else "1"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", "::regex:%s:%s" % (rx, idx)],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ins = []
for r in rows[2:]:
    if len(r) <= ismp or not r[ia].startswith("0x"):
        if ins:
            break      # next kernel section
        continue
    ins.append((r[isrc].strip(), int(r[iex]), int(r[ismp])))
tot_ex = sum(i[1] for i in ins); tot_s = sum(i[2] for i in ins)
print("instructions: %d static, %d executed (warp-level), %d samples" % (len(ins), tot_ex, tot_s))
if "--sass" in sys.argv:
    for k, (s, e, m) in enumerate(ins):
        print("%5d %6.2f%% %6.2f%%  %s" % (k, 100.0 * e / tot_ex, 100.0 * m / max(tot_s, 1), s))
else:
    # opcode histogram weighted by executions
    h = {}
    for s, e, m in ins:
        op = s.split()[0] if not s.startswith("@") else s.split()[1]
        op = op.split(".")[0]
        a = h.setdefault(op, [0, 0]); a[0] += e; a[1] += m
    for op, (e, m) in sorted(h.items(), key=lambda kv: -kv[1][0])[:40]:
        print("%-10s exec %6.2f%%  samples %6.2f%%" % (op, 100.0 * e / tot_ex, 100.0 * m / max(tot_s, 1)))
