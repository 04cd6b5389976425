"""dual-head launch vs the two separate convolutions, with the ring debug knobs"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import atvsnet_b200 as A
from atvsnet_b200.network import conv3d_raw, conv3d_dual


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
for cin in (8, 32):
    x = torch.randn(1, D, H, W, cin, device='cuda').to(torch.bfloat16)
    w1 = torch.randn(3, 3, 3, cin, 8, device='cuda') * 0.05
    w2 = torch.randn(3, 3, 3, cin, 16, device='cuda') * 0.05
    w32 = torch.randn(3, 3, 3, cin, 32, device='cuda') * 0.05
    st = torch.zeros((4, 128), dtype=torch.float64, device='cuda')
    for k in ('ATVS_RING_DEBUG', 'ATVS_RING_MINB', 'ATVS_RING_ZS'):
        os.environ.pop(k, None)
    r = dict(cin=cin)
    r['s1'] = timed(lambda: conv3d_raw(x, 'p1_%d' % cin, w1, 8, 1, False, True, st[0], raw_dtype=torch.float16))
    r['s2'] = timed(lambda: conv3d_raw(x, 'p2_%d' % cin, w2, 16, 2, False, True, st[1], raw_dtype=torch.float16))
    r['plain32'] = timed(lambda: conv3d_raw(x, 'p32_%d' % cin, w32, 32, 1, False, True, st[1], raw_dtype=torch.float16))
    r['dual'] = timed(lambda: conv3d_dual(x, 'p1_%d' % cin, w1, 'p2_%d' % cin, w2, st[2], st[3]))
    for dbg in (1, 2, 4, 6, 7):
        os.environ['ATVS_RING_DEBUG'] = str(dbg)
        r['dual_dbg%d' % dbg] = timed(lambda: conv3d_dual(x, 'p1_%d' % cin, w1, 'p2_%d' % cin, w2, st[2], st[3]))
    os.environ.pop('ATVS_RING_DEBUG')
    for zs in (8, 16, 32):
        os.environ['ATVS_RING_ZS'] = str(zs)
        r['dual_zs%d' % zs] = timed(lambda: conv3d_dual(x, 'p1_%d' % cin, w1, 'p2_%d' % cin, w2, st[2], st[3]))
    os.environ.pop('ATVS_RING_ZS')
    print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in r.items()}), flush=True)
