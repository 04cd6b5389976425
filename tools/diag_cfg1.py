import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import atvsnet_b200 as A
from oracle import schedule as osch
def mae(a,b,cams,D): return float(np.abs(a-b).mean())/((D-1)*float(cams[0,0,1,3,1]))
npy=lambda t: t.detach().float().cpu().numpy()
D=128
w = A.variables.synthetic_weights(); w.update(A.variables.synthetic_fem_weights()); w.update(A.variables.synthetic_refine_weights())
A.variables.load_weights(w)
for scene in ('2','0'):
    images, cams, _ = A.pipeline.load_example('tests/golden/example/'+scene, view_num=3)
    t=time.time()
    ref = osch.run_twoview(images, cams, D, w) if scene=='2' else osch.run_multiview(images, cams, D, w)
    print('scene', scene, 'oracle s', time.time()-t, flush=True)
    for prec in ('fp32','fp16','bf16'):
        A.FLAGS.precision=prec
        out = A.pipeline.run_example(torch.from_numpy(images).cuda(), torch.from_numpy(cams).cuda(), D)
        torch.cuda.synchronize()
        if scene=='2':
            print(prec, 'twoview refined_up', mae(npy(out['depth_refined_up']), ref['depth_refined_up'], cams, D),
                  'refined lo', mae(npy(out['depth_refined']), ref['depth_refined'], cams, D), flush=True)
        else:
            print(prec, 'stageII', mae(npy(out['depth']), ref['depth_agg_init'], cams, D), 'final up', mae(npy(out['depth_refined_up']), ref['depth_refined_up'], cams, D),
                  'views', [mae(npy(a), b, cams, D) for a,b in zip(out['depth_views'], ref['depth_views'])], flush=True)
    A.FLAGS.precision='fp16'
