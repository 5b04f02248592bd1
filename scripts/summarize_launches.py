"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count,
total and share of device time."""
import csv, sys, collections
for path in sys.argv[1:]:
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(lines); hdr = next(r)
    ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for row in r:
        name = row[ik].split('(')[0].replace('<unnamed>::', '').replace('void ', '')
        t = float(row[iv].replace(',', ''))
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
    tot = sum(a[1] for a in agg.values())
    print("## %s  (total %.3f ms over %d launches)" % (path, tot / 1e6, sum(a[0] for a in agg.values())))
    print("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.3f | %.1f | %.1f%% |" % (k, n, t / 1e6, t / n / 1e3, 100 * t / tot))
    print()
