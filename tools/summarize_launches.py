"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: python tools/summarize_launches.py f.csv [steps]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
d = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("dcgp::", "")
    d[name][0] += 1; d[name][1] += v
tot = sum(v[1] for v in d.values())
print("total %.3f ms per step over %d steps" % (tot / 1e6 / steps, steps))
for k, v in sorted(d.items(), key=lambda x: -x[1][1])[:45]:
    print("%8.3f ms %6.1f  %s" % (v[1] / 1e6 / steps, v[0] / steps, k[:90]))
