"""one process: the dominant 32->8 ring layer at cfg2 under launch knobs x role knock-outs (ATVS_RING_DEBUG bits:
1 no loads, 2 no MMAs, 4 no stores).   python tools/ring32_sweep.py [cin cout]"""
import sys, os, json, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import atvsnet_b200 as A
from atvsnet_b200.network import conv3d_raw
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from balanced_probe import timed  # noqa

cin, cout = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32, 8)
D, H, W = 128, 128, 160
KEYS = ('ATVS_RING_R', 'ATVS_RING_PF', 'ATVS_RING_MINB', 'ATVS_RING_DEBUG', 'ATVS_RING_BALANCED', 'ATVS_RING_CTAS')
x = torch.randn(1, D, H, W, cin, device='cuda').to(torch.float16)
w = torch.randn(3, 3, 3, cin, cout, device='cuda') * 0.05
stats = torch.zeros(2 * cout, dtype=torch.float64, device='cuda')
base = [{}, {'ATVS_RING_MINB': 1}, {'ATVS_RING_MINB': 1, 'ATVS_RING_R': 4, 'ATVS_RING_PF': 3},
        {'ATVS_RING_MINB': 1, 'ATVS_RING_R': 8, 'ATVS_RING_PF': 2}, {'ATVS_RING_MINB': 1, 'ATVS_RING_R': 8, 'ATVS_RING_PF': 6},
        {'ATVS_RING_R': 2, 'ATVS_RING_PF': 1}]
for kn in base:
    for dbg in (0, 5, 6, 3, 7):
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update({k: str(v) for k, v in kn.items()})
        os.environ['ATVS_RING_DEBUG'] = str(dbg)
        t = timed(lambda: conv3d_raw(x, 'r32', w, cout, 1, False, True, stats_buf=stats, raw_dtype=torch.float16))
        print(json.dumps(dict(cin=cin, cout=cout, knobs=kn, dbg=dbg, us=round(t, 1))), flush=True)
