import sys
rows = {}
for l in open(sys.argv[1]):
    f = l.split()
    if len(f) == 5 and f[0] in ('P', 'M', 'E'):
        rows.setdefault(f[0], []).append([int(v) for v in f[1:]])
n = {k: max(r[0] for r in v) + 1 for k, v in rows.items()}
last = {k: v[-n[k]:] for k, v in rows.items()}
t0 = min(r[1] for r in last['P'])
for role in 'PME':
    print(role, ' '.join('%d:(%d,%d,%d)' % (r[0], r[1] - t0, r[2] - t0, r[3] - t0) for r in last[role]))
