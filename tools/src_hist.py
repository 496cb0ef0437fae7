"""Per-source-line and per-opcode executed-instruction histogram of one kernel from `ncu --page source --csv`.
usage: src_hist.py <source.csv> [n_functions] [top]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
nf = float(sys.argv[2]) if len(sys.argv) > 2 else 1e6
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
per_line, src, opc = collections.Counter(), {}, collections.Counter()
cur_file, cur_line, ie, tot = None, None, None, 0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': ie = r.index('Instructions Executed'); continue
    if r[0] != '': cur_line = (cur_file, int(r[0])); src[cur_line] = r[1]; continue
    if r[2] in ('...', '-', ''): continue
    try: n = int(r[ie])
    except Exception: continue
    per_line[cur_line] += n; tot += n
    op = r[3].strip().split()
    o = op[1] if op[0].startswith('@') else op[0]
    opc[o.split('.')[0]] += n
print("warp instructions", tot, "per function", tot * 32 / nf)
for k, v in per_line.most_common(top):
    print(f"{v*32/nf:7.1f} {k[0]}:{k[1]}  {src[k][:100]}")
print()
print("  ".join(f"{k}:{v*32/nf:.0f}" for k, v in opc.most_common(30)))
