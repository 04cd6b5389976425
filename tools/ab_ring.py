"""A/B of a ring-kernel knob in steady state (D = 1024) and as a lone cfg2 launch:  python tools/ab_ring.py KNOB=VAL [KNOB=VAL ...]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import atvsnet_b200 as A
from atvsnet_b200.network import conv3d_raw


def timed(fn, iters=3):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); st = torch.cuda.Stream(); st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(iters): fn()
    torch.cuda.current_stream().wait_stream(st)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


variants = [{}] + [dict([a.split('=')]) for a in sys.argv[1:]]
for cin, cout, stride, tr, D in ((8, 8, 1, 0, 1024), (32, 8, 1, 0, 1024), (16, 16, 1, 0, 1024), (8, 16, 2, 0, 1024), (16, 8, 2, 1, 512),
                                 (8, 8, 1, 0, 128), (32, 8, 1, 0, 128)):
    shp = (1, D, 64, 80, cin) if tr else (1, D, 128, 160, cin)
    x = torch.randn(*shp, device='cuda').to(torch.float16)
    w = (torch.randn(3, 3, 3, cout, cin, device='cuda') if tr else torch.randn(3, 3, 3, cin, cout, device='cuda')) * 0.05
    for env in variants:
        for v in variants:
            for k in v: os.environ.pop(k, None)
        os.environ.update(env)
        stats = torch.zeros(2 * cout, dtype=torch.float64, device='cuda')
        us = timed(lambda: conv3d_raw(x, 'ab_%d_%d_%d_%d' % (cin, cout, stride, tr), w, cout, stride, bool(tr), True, stats_buf=stats, raw_dtype=torch.float16))
        nd = (2 * D if tr else D)
        print(json.dumps(dict(cin=cin, cout=cout, stride=stride, tr=tr, D=D, env=env, us=round(us, 1), per_cfg2=round(us * 128 / nd, 1))), flush=True)
    del x
