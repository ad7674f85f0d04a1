"""Timeline of the big kernels (>= thr us) per stream and the idle gaps between them: python tools/analyze_trace.py f.json.gz [thr_us]"""
import gzip, json, sys, collections
ev = json.load(gzip.open(sys.argv[1], "rt"))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 100.0
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
streams = collections.Counter(e["stream"] for e in ev)
busy = collections.defaultdict(float)
for e in ev:
    busy[e["stream"]] += e["dur"]
print("span %.2f ms; streams: %s" % ((ev[-1]["ts"] + ev[-1]["dur"] - t0) / 1e3, {k: (v, round(busy[k] / 1e3, 2)) for k, v in streams.items()}))
# union busy time over all streams
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in ev)
tot, cs, ce = 0.0, iv[0][0], iv[0][1]
gaps = []
for a, b in iv[1:]:
    if a > ce:
        tot += ce - cs
        gaps.append((ce - t0, a - ce))
        cs, ce = a, b
    else:
        ce = max(ce, b)
tot += ce - cs
print("device busy (any stream) %.2f ms, idle %.2f ms" % (tot / 1e3, sum(g[1] for g in gaps) / 1e3))
print("largest idle gaps (t_ms, gap_us):", [(round(a / 1e3, 2), round(g)) for a, g in sorted(gaps, key=lambda x: -x[1])[:12]])
print("--- kernels >= %g us" % thr)
for e in ev:
    if e["dur"] >= thr:
        print("%9.3f ms  %8.1f us  s%-3s %s" % ((e["ts"] - t0) / 1e3, e["dur"], e["stream"], e["name"][:70]))
