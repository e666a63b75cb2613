"""Per-kernel totals from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_summary.py launches.csv [skip_first_n_launches] > summary.csv"""
import csv, sys, collections, re
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1 + skip:]:
    v = float(r[i_val].replace(",", ""))
    u = r[i_unit]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms", "msecond") else v * 1e3
    name = re.sub(r"\(.*", "", r[i_name])
    name = name.replace("void ", "").replace("dq::", "")
    tot[name] += ms; cnt[name] += 1
total = sum(tot.values())
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches", "total_ms", "share"])
for k, v in tot.most_common():
    w.writerow([k, cnt[k], f"{v:.3f}", f"{v / total:.4f}"])
w.writerow(["TOTAL", sum(cnt.values()), f"{total:.3f}", "1.0"])
