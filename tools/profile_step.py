"""One cfg2 step bracketed by cudaProfilerStart/Stop (use with `ncu --profile-from-start off`)."""
import sys, os, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import atvsnet_b200 as A
ap = argparse.ArgumentParser()
ap.add_argument('--views', type=int, default=5)
ap.add_argument('--D', type=int, default=128)
ap.add_argument('--h', type=int, default=128)
ap.add_argument('--w', type=int, default=160)
ap.add_argument('--precision', default='fp16')
ap.add_argument('--no-siamese', action='store_true')
a = ap.parse_args()
A.FLAGS.precision = a.precision
A.variables.load_weights(A.variables.synthetic_weights())
cams = torch.from_numpy(A.synthetic.orbit_cams(a.views, a.h, a.w, a.D)[None]).cuda()
feats = torch.from_numpy(A.synthetic.smooth_features(a.views, a.h, a.w, 32, seed=0)[None]).cuda()
for _ in range(2):
    A.pipeline.run_multiview(feats, cams, a.D, siamese=not a.no_siamese)
torch.cuda.synchronize()
torch.cuda.profiler.start()
A.pipeline.run_multiview(feats, cams, a.D, siamese=not a.no_siamese)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
