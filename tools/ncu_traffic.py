"""per-launch DRAM traffic of every kernel in an `ncu --set full` report -> JSON (read by bench.py's roofline.traffic):
    python tools/ncu_traffic.py report.ncu-rep profiles/rNN_traffic.json"""
import csv, io, json, subprocess, sys


def num(v):
    return float(v.replace(',', '')) if v not in ('', 'n/a') else 0.0


SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0,
         'msecond': 1e3}
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h, units = r[0], r[1]
col = {k: i for i, k in enumerate(h)}
rows = []
for row in r[2:]:
    def get(k):
        i = col[k]
        return num(row[i]) * SCALE.get(units[i], 1.0)
    rows.append({"kernel": row[col['Kernel Name']], "grid": row[col['Grid Size']], "block": row[col['Block Size']],
                 "dram_bytes": get('dram__bytes_read.sum') + get('dram__bytes_write.sum'),
                 "dram_read_bytes": get('dram__bytes_read.sum'), "dram_write_bytes": get('dram__bytes_write.sum'),
                 "time_us_under_ncu": get('gpu__time_duration.sum'),
                 "tensor_pipe_active_pct": num(row[col['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']])
                 if 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active' in col else None,
                 "dram_throughput_pct": num(row[col['dram__throughput.avg.pct_of_peak_sustained_elapsed']])
                 if 'dram__throughput.avg.pct_of_peak_sustained_elapsed' in col else None})
json.dump(rows, open(sys.argv[2], 'w'), indent=1)
for x in rows:
    print('%-60s %8.1f MB  %8.1f us' % (x['kernel'][:60], x['dram_bytes'] / 1e6, x['time_us_under_ncu']))
