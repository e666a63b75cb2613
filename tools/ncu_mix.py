"""Instruction mix of one kernel from `ncu --page source --csv` output: python tools/ncu_mix.py file.csv"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
i_src, i_ex, i_s = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
body = [r for r in rows[2:] if len(r) > max(i_src, i_ex, i_s) and r[i_ex].isdigit()]
tot = 0; ops = collections.Counter(); samp = collections.Counter()
for r in body:
    n = int(r[i_ex])
    toks = r[i_src].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    ops[op] += n; samp[op] += int(r[i_s]); tot += n
print("sass lines", len(body), "total warp instr", tot)
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{op:10s} {n/tot*100:5.1f}%  ({n})  stall samples {samp[op]}")
