"""fixed (size-independent) cost of the convolution launches: graph-replayed timing of tiny volumes."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import atvsnet_b200 as A
from atvsnet_b200.network import conv3d_raw


def timed(fn, iters=20):
    for _ in range(3):
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


cases = [(8, 8, 1, 0), (32, 8, 1, 0), (16, 16, 1, 0), (8, 16, 2, 0), (32, 16, 2, 0), (16, 8, 2, 1), (64, 32, 2, 1), (64, 64, 1, 0)]
shapes = [(8, 16, 16), (16, 64, 80), (32, 64, 80), (64, 64, 80)]
for (cin, cout, stride, tr) in cases:
    for (D, H, W) in shapes:
        x = torch.randn(1, D, H, W, cin, device='cuda').to(torch.bfloat16)
        w = (torch.randn(3, 3, 3, cout, cin, device='cuda') if tr else torch.randn(3, 3, 3, cin, cout, device='cuda')) * 0.05
        stats = torch.zeros(2 * cout, dtype=torch.float64, device='cuda')
        key = 'fc_%d_%d_%d_%d' % (cin, cout, stride, tr)
        r = {}
        for minvox in ('65536', '1'):
            os.environ['ATVS_RING_MINVOX'] = minvox
            try:
                t = timed(lambda: conv3d_raw(x, key, w, cout, stride, tr, True, stats_buf=stats))
            except Exception as e:
                t = str(e)[:60]
            r['minvox' + minvox] = t
        print(json.dumps(dict(cin=cin, cout=cout, stride=stride, tr=tr, shape=[D, H, W], **r)), flush=True)
