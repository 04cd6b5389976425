"""one process: time conv layers under different env knobs (the C side reads getenv at every launch)"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import atvsnet_b200 as A
from atvsnet_b200.network import conv3d_raw


def timed(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(iters):
                fn()
    torch.cuda.current_stream().wait_stream(st)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


D, H, W = 128, 128, 160
knobs = [{}, {'ATVS_RING_R': '12', 'ATVS_RING_PF': '4'}, {'ATVS_RING_R': '12', 'ATVS_RING_PF': '6'},
         {'ATVS_RING_R': '12', 'ATVS_RING_PF': '8'}, {'ATVS_RING_R': '16', 'ATVS_RING_PF': '8'},
         {'ATVS_RING_R': '8', 'ATVS_RING_PF': '6'}, {'ATVS_RING_R': '16', 'ATVS_RING_PF': '8', 'ATVS_RING_MINB': '1'}]
for (cin, cout) in ((8, 8), (8, 16), (16, 8), (32, 8)):
    x = torch.randn(1, D, H, W, cin, device='cuda').to(torch.bfloat16)
    w = torch.randn(3, 3, 3, cin, cout, device='cuda') * 0.05
    stats = torch.zeros(2 * cout, dtype=torch.float64, device='cuda')
    for kn in knobs:
        for k in ('ATVS_RING_R', 'ATVS_RING_PF', 'ATVS_RING_MINB'):
            os.environ.pop(k, None)
        os.environ.update(kn)
        t = timed(lambda: conv3d_raw(x, 'ks%d_%d' % (cin, cout), w, cout, 1, False, True, stats_buf=stats))
        print(json.dumps(dict(cin=cin, cout=cout, knobs=kn, us=round(t, 1))), flush=True)
