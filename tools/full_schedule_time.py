"""time the whole four-stage example.py schedule (images in: FEM, stages I-IV) at cfg2, eager launches, per stage"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import atvsnet_b200 as A

nv, H, W, D = 5, 512, 640, 128
w = A.variables.synthetic_weights()
w.update(A.variables.synthetic_fem_weights())
w.update(A.variables.synthetic_refine_weights())
A.variables.load_weights(w)
rng = np.random.default_rng(0)
imgs = torch.from_numpy((127.5 + 50 * rng.standard_normal((1, nv, H, W, 3))).clip(0, 255).astype(np.float32)).cuda()
cams = torch.from_numpy(A.synthetic.orbit_cams(nv, H // 4, W // 4, D)[None]).cuda()


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return (time.time() - t) / n * 1e3, out


t_all, out = timed(lambda: A.pipeline.run_example_schedule(imgs, cams, D))
t_fem, feats = timed(lambda: A.fem.extract_features(imgs))
t_12, s12 = timed(lambda: A.pipeline.run_multiview(feats, cams, D, siamese=True, upsample=False))
ds, di = cams[:, 0, 1, 3, 0].contiguous(), cams[:, 0, 1, 3, 1].contiguous()
t_3, _ = timed(lambda: A.refine.TVSNet_refine(s12['depth'], s12['depth_views'][0], s12['prob_volume_agg'], s12['cost_volume_agg'],
                                               imgs, cams, D, ds, di, 1))
d = out['depth_refined_up']
print(json.dumps(dict(workload='cfg2 from images, all four stages, eager', ms_total=t_all, ms_fem_5_views=t_fem, ms_stage_1_2=t_12,
                      ms_stage3_per_source=t_3, depth_finite=bool(torch.isfinite(d).all()), shape=list(d.shape),
                      peak_mem_gb=torch.cuda.max_memory_allocated() / 1e9)))
