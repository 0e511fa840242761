"""Reads `ncu -i X.ncu-rep --page source --csv` (SASS view) from stdin or a file: instruction mix by opcode and the hottest
instructions by stall samples."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ci = {n: i for i, n in enumerate(hdr)}
ie, smp, src = ci["Instructions Executed"], ci["# Samples"], ci["Source"]
tot = tots = 0
byop = collections.Counter(); byops = collections.Counter()
lines = []
for r in rows[h + 1:]:
    try:
        v = float(r[ie]); s = float(r[smp])
    except Exception:
        continue
    op = r[src].split()[0] if not r[src].startswith("@") else r[src].split()[1]
    op = op.split(".")[0]
    byop[op] += v; byops[op] += s; tot += v; tots += s
    lines.append((s, v, r[src]))
print("total warp instructions %.3e, samples %d" % (tot, tots))
for op, v in byop.most_common(22):
    print("%-10s %6.2f%% instr  %6.2f%% samples" % (op, 100 * v / tot, 100 * byops[op] / max(tots, 1)))
print("--- hottest by samples")
lines.sort(reverse=True)
for s, v, t in lines[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%5.2f%% smp %5.2f%% ins  %s" % (100 * s / max(tots, 1), 100 * v / tot, t[:110]))
