"""Balanced plane ranges (ATVS_RING_BALANCED, conv_ring.cu UnitIter) against fixed z segments: every ring layer of the
cfg2 step alone (CUDA-graph replay of 10 launches) and the whole step, one process (the C side reads getenv per launch).

    python tools/balanced_probe.py [kernels|step|all]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import atvsnet_b200 as A
from atvsnet_b200.network import conv3d_raw

KEYS = ('ATVS_RING_BALANCED', 'ATVS_RING_CTAS', 'ATVS_RING_MINPLANES', 'ATVS_RING_CTAS_32_8', 'ATVS_RING_CTAS_8_8',
        'ATVS_RING_CTAS_8_16', 'ATVS_RING_CTAS_16_16', 'ATVS_RING_CTAS_8_1', 'ATVS_S2_BALANCED', 'ATVS_DRING_BALANCED',
        'ATVS_S2_CTAS', 'ATVS_DRING_CTAS')


def setenv(kn):
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in kn.items()})


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
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / iters)
    return best


def kernels():
    layers = [(32, 8, 1, False, (128, 128, 160)), (8, 8, 1, False, (128, 128, 160)), (8, 16, 1, False, (128, 128, 160)),
              (16, 16, 1, False, (64, 64, 80)), (8, 16, 2, False, (128, 128, 160)), (16, 8, 2, True, (64, 64, 80)),
              (32, 16, 2, True, (32, 32, 40))]
    variants = [{'ATVS_RING_BALANCED': 0, 'ATVS_S2_BALANCED': 0, 'ATVS_DRING_BALANCED': 0}, {},
                {'ATVS_RING_CTAS': 148, 'ATVS_S2_CTAS': 148, 'ATVS_DRING_CTAS': 148},
                {'ATVS_RING_CTAS': 222, 'ATVS_S2_CTAS': 222, 'ATVS_DRING_CTAS': 222},
                {'ATVS_RING_MINPLANES': 6}, {'ATVS_RING_MINPLANES': 20}]
    for cin, cout, stride, tr, (D, H, W) in layers:
        x = torch.randn(1, D, H, W, cin, device='cuda').to(torch.float16)
        w = (torch.randn(3, 3, 3, cout, cin, device='cuda') if tr else torch.randn(3, 3, 3, cin, cout, device='cuda')) * 0.05
        stats = torch.zeros(2 * cout, dtype=torch.float64, device='cuda')
        for kn in variants:
            setenv(kn)
            t = timed(lambda: conv3d_raw(x, 'bp%d_%d_%d_%d' % (cin, cout, stride, tr), w, cout, stride, tr, True, stats_buf=stats, raw_dtype=torch.float16))
            print(json.dumps(dict(cin=cin, cout=cout, stride=stride, transposed=tr, shape=[D, H, W], knobs=kn,
                                  us=round(t, 1))), flush=True)
    setenv({})


def step():
    nv, h, w, D = 5, 128, 160, 128
    A.variables.load_weights(A.variables.synthetic_weights())
    cams = torch.from_numpy(A.synthetic.orbit_cams(nv, h, w, D)[None]).cuda()
    feats = torch.from_numpy(A.synthetic.smooth_features(nv, h, w, 32, seed=0)[None]).cuda()
    fn = lambda: A.pipeline.run_multiview(feats, cams, D, siamese=True)['depth_up']
    variants = [{'ATVS_RING_BALANCED': 0, 'ATVS_S2_BALANCED': 0, 'ATVS_DRING_BALANCED': 0}, {},
                {'ATVS_RING_CTAS': 148}, {'ATVS_RING_CTAS': 222}, {'ATVS_RING_CTAS_8_8': 148},
                {'ATVS_RING_CTAS_8_8': 200}, {'ATVS_RING_CTAS_32_8': 222}, {'ATVS_RING_CTAS_16_16': 64},
                {'ATVS_RING_MINPLANES': 20}, {'ATVS_RING_BALANCED': 0}, {'ATVS_S2_BALANCED': 0}, {'ATVS_DRING_BALANCED': 0}]
    if len(sys.argv) > 2:
        variants = [json.loads(a) for a in sys.argv[2:]]
    for kn in variants:
        setenv(kn)
        t = timed(fn, iters=8)
        print(json.dumps(dict(step_ms=round(t / 1e3, 3), knobs=kn)), flush=True)
    setenv({})


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if what in ('kernels', 'all'):
        kernels()
    if what in ('step', 'all'):
        step()
