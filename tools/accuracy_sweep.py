"""bf16 tensor-core path vs fp32 CUDA-core path (which tracks the CPU oracle to <1e-3 of the
range, tests/test_gpu_parity.py) across volume sizes: depth MAE as a fraction of the range."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import atvsnet_b200 as A

def run(D, h, w, nv, prec, gain, raw='f16'):
    A.FLAGS.precision = prec
    A.FLAGS.raw_dtype = raw
    cams = torch.from_numpy(A.synthetic.orbit_cams(nv, h, w, D)[None]).cuda()
    feats = torch.from_numpy(A.synthetic.smooth_features(nv, h, w, 32, seed=3)[None]).cuda()
    torch.cuda.synchronize(); t = time.time()
    out = A.pipeline.run_multiview(feats, cams, D, siamese=False)
    torch.cuda.synchronize(); dt = time.time() - t
    return out, dt, float(cams[0, 0, 1, 3, 1]) * (D - 1)

for gain in (2.0, 4.0):
    A.variables.load_weights(A.variables.synthetic_weights(seed=11, logit_gain=gain))
    for (D, h, w, nv) in ((16, 16, 24, 3), (32, 32, 48, 3), (64, 64, 80, 3), (128, 128, 160, 3)):
        o32, t32, rng = run(D, h, w, nv, 'fp32', gain)
        p = torch.softmax(-o32['prob_volume_agg'], dim=1).max(dim=1).values.mean().item()
        for raw in ('f32', 'f16'):
            o16, t16, _ = run(D, h, w, nv, 'bf16', gain, raw)
            mae = (o16['depth_up'] - o32['depth_up']).abs().mean().item() / rng
            mx = (o16['depth_up'] - o32['depth_up']).abs().max().item() / rng
            print(json.dumps(dict(gain=gain, D=D, h=h, w=w, raw=raw, mae_over_range=mae, max_over_range=mx, mean_peak_prob=p,
                                  t_fp32=t32, t_bf16=t16)), flush=True)
