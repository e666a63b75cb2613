"""Shared-memory wavefronts per SASS line (actual vs ideal) from an .ncu-rep: python tools/ncu_smem.py rep kernel-regex"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[2]],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
b = blocks[0]
h = b["rows"][0]; ix = {k: i for i, k in enumerate(h)}
body = [r for r in b["rows"][1:] if len(r) > ix["# Samples"] and r[ix["Instructions Executed"]].isdigit()]
sh = [(int(r[ix["L1 Wavefronts Shared"]] or 0), int(r[ix["L1 Wavefronts Shared Ideal"]] or 0), int(r[ix["Instructions Executed"]]), i,
       r[ix["Source"]].strip()) for i, r in enumerate(body) if (r[ix["L1 Wavefronts Shared"]] or "0") != "0"]
W = sum(s[0] for s in sh); I = sum(s[1] for s in sh)
print(b["name"], "shared wavefronts", W, "ideal", I, "excess", W - I)
agg = {}
for w, i, n, idx, s in sh:
    op = s.split()[1] if s.startswith("@") else s.split()[0]
    a = agg.setdefault(op, [0, 0, 0]); a[0] += w; a[1] += i; a[2] += n
for op, a in sorted(agg.items(), key=lambda kv: -kv[1][0]): print(f"  {op:12s} wavefronts {a[0]:>12d} ideal {a[1]:>12d} instr {a[2]:>11d}")
sh.sort(reverse=True)
for s in sh[: int(sys.argv[3]) if len(sys.argv) > 3 else 30]: print("   ", s)
