"""summarise the ATVS_RING_TRACE dump of conv_ring.cu (CTA 0): python tools/ring_trace.py file"""
import sys
for fn in sys.argv[1:]:
    last = {}
    for l in open(fn):
        if l[:2] in ('P ', 'M ', 'E '):
            f = l.split()
            last.setdefault(f[0], []).append([int(v) for v in f[1:]])
    print(fn)
    t0 = min(r[1] for r in last['P'][-48:])
    for role in 'PME':
        rows = last[role][-48:]
        print(role, ' '.join('%d:(%d,%d,%d)' % (r[0], r[1] - t0, r[2] - t0, r[3] - t0) for r in rows[:3]), '...')
        print('   ', ' '.join('%d:(%d,%d,%d)' % (r[0], r[1] - t0, r[2] - t0, r[3] - t0) for r in rows[24:30]))
        n = len(rows)
        d = [rows[i + 1][1] - rows[i][1] for i in range(8, min(40, n - 1))]
        ab = [r[2] - r[1] for r in rows[8:40]]
        bc = [r[3] - r[2] for r in rows[8:40]]
        print('   mean period %.0f clk, A->B %.0f, B->C %.0f' % (sum(d) / len(d), sum(ab) / len(ab), sum(bc) / len(bc)))
