"""Steady-state knock-outs of the ring kernels: a volume 8x deeper than cfg2 (D = 1024: launch / ramp / drain costs are
<3 % of the launch), ATVS_RING_DEBUG bits 1 no loads, 2 no MMAs, 4 no stores, 8 no accumulator zeroing.
    python tools/steady_probe.py"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import atvsnet_b200 as A
from atvsnet_b200.network import conv3d_raw


def timed(fn, iters=3):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(iters):
                fn()
    torch.cuda.current_stream().wait_stream(st)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


D, H, W = 1024, 128, 160
cases = [(8, 8, 1, 0), (32, 8, 1, 0), (8, 16, 2, 0), (16, 8, 2, 1)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]]
for cin, cout, stride, tr in cases:
    shp = (1, D // 2, H // 2, W // 2, cin) if tr else (1, D, H, W, cin)
    x = torch.randn(*shp, device='cuda').to(torch.float16)
    w = (torch.randn(3, 3, 3, cout, cin, device='cuda') if tr else torch.randn(3, 3, 3, cin, cout, device='cuda')) * 0.05
    key = 'sp_%d_%d_%d_%d' % (cin, cout, stride, tr)
    for dbg in (0, 1, 2, 4, 3, 5, 6, 7):
        os.environ['ATVS_RING_DEBUG'] = str(dbg)
        stats = torch.zeros(2 * cout, dtype=torch.float64, device='cuda')
        us = timed(lambda: conv3d_raw(x, key, w, cout, stride, bool(tr), True, stats_buf=stats, raw_dtype=torch.float16))
        print(json.dumps(dict(cin=cin, cout=cout, stride=stride, tr=tr, dbg=dbg, us=round(us, 1), us_per_cfg2=round(us / 8, 1))), flush=True)
    os.environ['ATVS_RING_DEBUG'] = '0'
    del x
