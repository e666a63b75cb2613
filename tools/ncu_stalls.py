"""Per-kernel stall breakdown + top stalled instructions from an .ncu-rep: python tools/ncu_stalls.py rep [kernel-regex] [stall]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "."
which = sys.argv[3] if len(sys.argv) > 3 else "stall_long_sb"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# split per kernel: a "Kernel Name" row starts each block
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
seen = set()
for b in blocks:
    if b["name"] in seen: continue
    seen.add(b["name"])
    h = b["rows"][0]
    ix = {k: i for i, k in enumerate(h)}
    st = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    body = [r for r in b["rows"][1:] if len(r) > ix["# Samples"] and r[ix["Instructions Executed"]].isdigit()]
    tot = sum(int(r[ix["# Samples"]]) for r in body)
    inst = sum(int(r[ix["Instructions Executed"]]) for r in body)
    print("==", b["name"], "samples", tot, "warp-instr", inst)
    agg = sorted(((sum(int(r[ix[k]] or 0) for r in body) / max(tot, 1), k) for k in st), reverse=True)
    print("  ", ", ".join(f"{k[6:]} {v:.3f}" for v, k in agg[:9]))
    top = sorted(((int(r[ix[which]] or 0), r[ix["Source"]].strip(), r[ix["Instructions Executed"]]) for r in body), reverse=True)
    for x in top[:10]: print("     ", x)
