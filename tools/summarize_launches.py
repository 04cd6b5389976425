"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, gi, vi, mi = hdr.index('Kernel Name'), hdr.index('Grid Size'), hdr.index('Metric Value'), hdr.index('Metric Name')
ui = hdr.index('Metric Unit')
agg = collections.OrderedDict()
tot = 0.0
for r in rows[1:]:
    if r[mi] != 'gpu__time_duration.sum':
        continue
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] in ('ns', 'nsecond') else (v * 1e3 if r[ui].startswith('ms') else v)   # -> us
    name = re.sub(r'\(.*', '', r[ki])
    name = re.sub(r'^void |\(anonymous namespace\)::', '', name)
    key = name[:70]
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print('total %.1f us over %d launches' % (tot, sum(a[0] for a in agg.values())))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%9.1f us  %5.1f%%  n=%4d  avg %8.1f us  %s' % (t, 100 * t / tot, n, t / n, k))
