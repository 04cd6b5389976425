"""deconv plane-ring kernel alone on the cfg2 shapes: time vs z-segment length / CTAs per SM / ring depth."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import atvsnet_b200 as A
from atvsnet_b200.network import conv3d_raw

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for (cin, cout, shape) in ((16, 8, (1, 64, 64, 80)), (32, 16, (1, 32, 32, 40))):
    x = torch.randn(shape + (cin,), device='cuda').half()
    w = (torch.randn(3, 3, 3, cout, cin, device='cuda') * 0.05)
    stats = torch.zeros(128, dtype=torch.float64, device='cuda')
    fn = lambda: conv3d_raw(x, 'probe%d' % cin, w, cout, 2, True, True, stats, raw_dtype=torch.float16)
    os.environ['ATVS_NO_DECONV_RING'] = '1'
    print(cin, cout, shape, 'old kernel: %.1f us' % timeit(fn), flush=True)
    del os.environ['ATVS_NO_DECONV_RING']
    for zs in (2, 4, 8, 11, 16, 22, 32, 64):
        if zs > shape[1]: continue
        for minb in ((2, 1) if cout == 8 else (1,)):
            os.environ['ATVS_DRING_ZS'] = str(zs)
            os.environ['ATVS_DRING_MINB'] = str(minb)
            print('  zs=%2d minb=%d: %.1f us' % (zs, minb, timeit(fn)), flush=True)
    os.environ.pop('ATVS_DRING_ZS'); os.environ.pop('ATVS_DRING_MINB')
    for r in (2, 3, 4, 6, 8):
        os.environ['ATVS_DRING_R'] = str(r)
        print('  default zs, R=%d: %.1f us' % (r, timeit(fn)), flush=True)
    os.environ.pop('ATVS_DRING_R')
