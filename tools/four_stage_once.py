"""one four-stage schedule (example.py:144-181) at cfg2 between cudaProfilerStart/Stop (ncu --profile-from-start off);
features precomputed so that the list shows stages I-IV without the FEM"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import atvsnet_b200 as A
import bench
w = A.variables.synthetic_weights(); w.update(A.variables.synthetic_fem_weights()); w.update(A.variables.synthetic_refine_weights())
A.variables.load_weights(w)
feats, cams, D = bench.make_inputs('cfg2', 0)
rng = np.random.default_rng(1000)
imgs = torch.from_numpy((127.5 + 50.0 * rng.standard_normal((1, 5, 512, 640, 3))).clip(0, 255).astype(np.float32)).cuda()
f, c = torch.from_numpy(feats).cuda(), torch.from_numpy(cams).cuda()
for _ in range(2):
    A.pipeline.run_example_schedule(imgs, c, D, features=f)
torch.cuda.synchronize()
torch.cuda.profiler.start()
A.pipeline.run_example_schedule(imgs, c, D, features=f)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
