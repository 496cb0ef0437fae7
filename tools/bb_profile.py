"""Basic-block view of `ncu --page source --csv --print-source sass` output: runs of consecutive SASS
instructions with the same executed count.  usage: bb_profile.py <sass.csv> [n_functions]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
nf = float(sys.argv[2]) if len(sys.argv) > 2 else 1e6
hdr = None
for i, r in enumerate(rows):
    if r and r[0] in ('Address', '#', 'Line No') or (r and 'Instructions Executed' in r):
        hdr = i; break
h = rows[hdr]
ie = h.index('Instructions Executed'); isrc = h.index('Source')
blocks = []
cur = None
for r in rows[hdr + 1:]:
    if len(r) <= ie: continue
    try: n = int(r[ie])
    except Exception: continue
    ins = r[isrc].strip()
    if cur and cur[0] == n: cur[1] += 1; cur[3].append(ins)
    else:
        cur = [n, 1, ins, [ins]]; blocks.append(cur)
tot = sum(b[0] * b[1] for b in blocks)
print("total warp instr", tot, "per function", tot * 32 / nf)
for b in blocks:
    w = b[0] * b[1] * 32 / nf
    if w >= 3.0:
        ops = {}
        for x in b[3]:
            p = x.split()
            o = (p[1] if p[0].startswith('@') else p[0]).split('.')[0]
            ops[o] = ops.get(o, 0) + 1
        top = " ".join(f"{k}{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:7])
        print(f"{w:7.1f}/fn  execs/warp-fn {b[0]*32/nf:5.2f}  n_instr {b[1]:4d}  {top}")
