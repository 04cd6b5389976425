"""top stall-sample instructions of an ncu report (source page): python tools/ncu_hot.py rep [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))[2:]
rows = [(int(x[2]), int(x[5]), i, x[1].strip()) for i, x in enumerate(r) if len(x) > 5 and x[2].isdigit()]
tot = sum(a for a, _, _, _ in rows)
print('total samples', tot, 'instructions', len(rows))
for a, e, i, s in sorted(rows, reverse=True)[:n]:
    print('%6d %5.1f%% exec=%8d  #%4d  %s' % (a, 100.0 * a / tot, e, i, s[:100]))
