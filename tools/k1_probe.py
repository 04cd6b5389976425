"""run K1 (bf16 warped-only) / K4 a few times at one size, for ncu: python tools/k1_probe.py h w D"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import atvsnet_b200 as A
h, w, D = (int(v) for v in sys.argv[1:4])
cams = torch.from_numpy(A.synthetic.orbit_cams(2, h, w, D)[None]).cuda()
feats = torch.from_numpy(A.synthetic.smooth_features(2, h, w, 32, seed=1)[None]).cuda()
ds, di = cams[:, 0, 1, 3, 0].contiguous(), cams[:, 0, 1, 3, 1].contiguous()
vol = torch.randn(1, D, h, w, device='cuda')
for _ in range(5):
    out = A.build_cost_volume(feats[:, 0], feats[:, 1], cams, D, ds, di, 0, 1, mode='warped_only', out_dtype=torch.bfloat16)
    A.prob2depth(vol, D, ds, di)
    if h * w <= 512 * 640:
        A.prob2depth_upsample(vol, D, ds, di)
torch.cuda.synchronize()
