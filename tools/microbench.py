"""Kernel microbench sweep (BASELINE.json configs[4]): K1 (fused homography + bilinear + cost slice) and
K4 (soft-argmin, plain and fused x4) against the HBM roofline, CUDA-graph replay + CUDA events.

    python tools/microbench.py [--quick] [--out gpurun_out/microbench.json]

H x W is taken DIRECTLY as the plane size; algorithmic bytes:
  K1 : 4*h*w*F (source) [+ 4*h*w*F ref in CONCAT mode] + s*D*h*w*C_out         (s = 2 bf16, 4 fp32)
  K4 : 4*D*h*w read + 4*h'*w' write (h' = up*h)
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import atvsnet_b200 as A

ap = argparse.ArgumentParser()
ap.add_argument('--quick', action='store_true')
ap.add_argument('--out', default=None)
ap.add_argument('--only', default=None, help='k1|k4|k4up')
a = ap.parse_args()
pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json'))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0}
PEAK = pk['hbm_gbs']
F = 32


def timed(fn, iters):
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
    return e0.elapsed_time(e1) * 1e-3 / iters


sizes = [(256, 320), (512, 640), (1024, 1280), (1080, 1920)]
depths = [48, 96, 128, 192, 256]
if a.quick:
    sizes, depths = [(128, 160), (512, 640)], [128]
rows = []
for (h, w) in sizes:
    cams = torch.from_numpy(A.synthetic.orbit_cams(2, h, w, 256)[None]).cuda()
    feats = torch.from_numpy(A.synthetic.smooth_features(2, h, w, F, seed=1)[None]).cuda()
    for D in depths:
        cams_d = torch.from_numpy(A.synthetic.orbit_cams(2, h, w, D)[None]).cuda()
        ds, di = cams_d[:, 0, 1, 3, 0].contiguous(), cams_d[:, 0, 1, 3, 1].contiguous()
        V = D * h * w
        if V * F * 2 > 40e9:
            continue
        iters = max(2, min(20, int(2e9 // (V * F * 2))))
        if a.only in (None, 'k1'):
            for mode, dt, s, cout in (('warped_only', torch.bfloat16, 2, F), ('concat', torch.float32, 4, 2 * F)):
                if V * cout * s > 40e9:
                    continue
                out = {}
                def run():
                    out['v'] = A.build_cost_volume(feats[:, 0], feats[:, 1], cams_d, D, ds, di, 0, 1, mode=mode, out_dtype=dt)
                t = timed(run, iters)
                nbytes = 4 * h * w * F * (2 if mode == 'concat' else 1) + s * V * cout
                rows.append(dict(kernel='K1 ' + mode + (' bf16' if s == 2 else ' fp32'), h=h, w=w, D=D, us=t * 1e6,
                                 gbs=nbytes / t / 1e9, frac=nbytes / t / 1e9 / PEAK, bytes=nbytes))
                print(json.dumps(rows[-1]), flush=True)
                del out
        vol = torch.randn(1, D, h, w, device='cuda')
        if a.only in (None, 'k4'):
            t = timed(lambda: A.prob2depth(vol, D, ds, di), iters)
            nbytes = 4 * V + 4 * h * w
            rows.append(dict(kernel='K4 soft-argmin', h=h, w=w, D=D, us=t * 1e6, gbs=nbytes / t / 1e9,
                             frac=nbytes / t / 1e9 / PEAK, bytes=nbytes))
            print(json.dumps(rows[-1]), flush=True)
        if a.only in (None, 'k4up') and h * w <= 512 * 640:
            t = timed(lambda: A.prob2depth_upsample(vol, D, ds, di), iters)
            nbytes = 2 * 4 * V + 4 * h * w * 17     # the op also returns the low-res estimate (second pass over the volume)
            rows.append(dict(kernel='K4 x4-upsample + low-res', h=h, w=w, D=D, us=t * 1e6, gbs=nbytes / t / 1e9,
                             frac=nbytes / t / 1e9 / PEAK, bytes=nbytes))
            print(json.dumps(rows[-1]), flush=True)
        del vol
if a.out:
    json.dump(dict(peak_hbm_gbs=PEAK, rows=rows), open(a.out, 'w'), indent=1)
