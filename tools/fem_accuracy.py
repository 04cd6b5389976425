"""depth effect of running the FEM's stride-1 convolutions on the tensor cores (fp16 operands): images -> features
(fp32 CUDA-core FEM | tensor-core FEM) -> stages I + II (fp16 path) -> depth map, against the fp32 CUDA path end to end."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import atvsnet_b200 as A
w = A.variables.synthetic_weights(); w.update(A.variables.synthetic_fem_weights())
A.variables.load_weights(w)
res = {}
for name, shape in (('synthetic 3 views 256x320 D=64', (3, 256, 320, 64)), ('example/0 3 views', None)):
    if shape is not None:
        nv, H, W, D = shape
        rng = np.random.default_rng(5)
        base = rng.standard_normal((1, 1, H // 8, W // 8, 3)).astype(np.float32)
        imgs = torch.nn.functional.interpolate(torch.from_numpy(base[0, 0]).permute(2, 0, 1)[None], size=(H, W), mode='bicubic')[0].permute(1, 2, 0).numpy()
        imgs = np.stack([np.roll(imgs, 3 * i, axis=1) for i in range(nv)])[None]
        imgs = (127.5 + 60 * imgs + 5 * rng.standard_normal(imgs.shape)).clip(0, 255).astype(np.float32)
        cams = A.synthetic.orbit_cams(nv, H // 4, W // 4, D)[None]
    else:
        imgs, cams, _ = A.pipeline.load_example('tests/golden/example/0', view_num=3)
        D = 128
    ti, tc = torch.from_numpy(imgs).cuda(), torch.from_numpy(cams).cuda()
    rng_ = (D - 1) * float(cams[0, 0, 1, 3, 1])
    A.FLAGS.precision = 'fp32'
    f32 = A.fem.extract_features(ti)
    d32 = A.pipeline.run_multiview(f32, tc, D, siamese=False)['depth_up']
    A.FLAGS.precision = 'fp16'
    A.FLAGS.fem_tensor = False
    f_cc = A.fem.extract_features(ti)
    A.FLAGS.fem_tensor = True
    f_tc = A.fem.extract_features(ti)
    d_a = A.pipeline.run_multiview(f32, tc, D, siamese=False)['depth_up']        # fp32 FEM, fp16 hot path
    d_b = A.pipeline.run_multiview(f_tc, tc, D, siamese=False)['depth_up']       # tensor FEM, fp16 hot path
    rel = lambda a, b: float((a - b).abs().mean() / b.abs().mean())
    res[name] = dict(feat_mean_rel_err_tensor_fem=rel(f_tc, f32), feat_mean_rel_err_fp32_fem_rerun=rel(f_cc, f32),
                     depth_mae_fp16_hotpath_only=float((d_a - d32).abs().mean()) / rng_,
                     depth_mae_tensor_fem_plus_fp16_hotpath=float((d_b - d32).abs().mean()) / rng_)
    print(name, json.dumps(res[name]), flush=True)
