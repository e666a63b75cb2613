"""SASS lines of the hottest loop (by executed count): python tools/ncu_hot.py src.csv [count_rank]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
i_src, i_ex, i_s = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
body = [r for r in rows[2:] if len(r) > i_s and r[i_ex].isdigit()]
c = collections.Counter(int(r[i_ex]) for r in body if int(r[i_ex]) > 0)
top = sorted(c.items(), key=lambda kv: -kv[0] * kv[1])
print("exec-count groups (count, lines):", top[:6])
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cnt = top[rank][0]
ops = collections.Counter()
for r in body:
    if int(r[i_ex]) == cnt:
        toks = r[i_src].split()
        op = toks[1] if toks[0].startswith('@') else toks[0]
        ops[op.split('.')[0]] += 1
        if len(sys.argv) > 3: print(r[i_src].strip()[:90], r[i_s])
print(sorted(ops.items(), key=lambda kv: -kv[1]))
