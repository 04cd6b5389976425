import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import atvsnet_b200 as A
from atvsnet_b200 import network as N
D, h, w, nv = 64, 64, 80, 3
A.variables.load_weights(A.variables.synthetic_weights(seed=11, logit_gain=2.0))
cams = torch.from_numpy(A.synthetic.orbit_cams(nv, h, w, D)[None]).cuda()
feats = torch.from_numpy(A.synthetic.smooth_features(nv, h, w, 32, seed=3)[None]).cuda()
for streams in (1, 4):
    for split in (False, True):
        A.pipeline.CONCURRENT_PASSES = streams
        A.pipeline.SPLIT_COST_VOLUME = split
        A.FLAGS.precision = A.flags.DEFAULT_PRECISION
        A.variables.packed_cache().clear()
        for rep in range(3):
            out = A.pipeline.run_multiview(feats, cams, D, siamese=False)
            torch.cuda.synchronize()
            bad = {k: int((~torch.isfinite(v.float())).sum()) for k, v in out.items() if torch.is_tensor(v)}
            print(json.dumps(dict(streams=streams, split=split, rep=rep, bad=bad)))
# per-layer hunt in a single pass
A.pipeline.CONCURRENT_PASSES = 1
A.pipeline.SPLIT_COST_VOLUME = True
ds, di = cams[:, 0, 1, 3, 0].contiguous(), cams[:, 0, 1, 3, 1].contiguous()
cv = A.pipeline._cost_volume(feats[:, 0], feats[:, 1], cams, D, ds, di, 0, 1)
print('warped finite', bool(torch.isfinite(cv.warped.float()).all()))
t = A.StackedUNet_prob({'data': cv}, outputs=tuple(n for n in A.StackedUNet_prob({'data': cv}).nodes if n != 'data'))
t.get_output()
for n, node in t.nodes.items():
    v = node.value
    if torch.is_tensor(v):
        nb = int((~torch.isfinite(v.float())).sum())
        if nb:
            print('first bad layer', n, nb, tuple(v.shape)); break
else:
    print('no bad layer in single pass')
