"""GPU, BASELINE.json configs[0]: atvsnet/example.py on the bundled scenes with 3 views - the REAL images and cameras
(tests/golden/example/, copied from the reference's example/ by make_example_fixtures.py), seeded synthetic weights
under the checkpoint names (the released model.zip is not available offline), CUDA path against the
same schedule composed from the CPU oracles (oracle/schedule.py):

  example/0  3 views, 960x640 images -> 160x240 features, D = 128: the four-stage multi-view schedule
             (example.py:144-191), images in, refinement included;
  example/2  2 views, 640x480: the two-view network TVSNet + prob2depth_upsample (example.py:219-272).

Each test runs ~1-2 minutes of CPU oracle next to the GPU run (pytest -m "gpu and not slow" skips them)."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.slow]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, 'tests', 'golden', 'example')


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def npy(t):
    return t.detach().float().cpu().numpy()


@pytest.fixture(scope='module')
def A():
    import atvsnet_b200 as A_
    return A_


def all_weights(A):
    w = A.variables.synthetic_weights()          # bench.py's CRM / AAM / output-conv weights (logit gain 4)
    w.update(A.variables.synthetic_fem_weights())
    w.update(A.variables.synthetic_refine_weights())
    return w


def mae_over_range(a, b, cams, D):
    return float(np.abs(a - b).mean()) / ((D - 1) * float(cams[0, 0, 1, 3, 1]))


def _final_from_stage2(A, s2, images, cams, D):
    """stages III + IV of example.py:163-181 on the CUDA path, fed with GIVEN stage-II results (dict of numpy arrays)."""
    from atvsnet_b200 import network as N
    ds, di = cu(cams[:, 0, 1, 3, 0]), cu(cams[:, 0, 1, 3, 1])
    rcs = []
    for n, v in enumerate(range(1, cams.shape[1])):
        _, rc = A.TVSNet_refine(cu(s2['depth_agg_init']), cu(s2['depth_views'][n]), cu(s2['prob_volume_agg']),
                                cu(s2['cost_volume_agg']), cu(images), cu(cams), D, ds, di, v)
        rcs.append(rc)
    cost_ref = N.attention_aggregation(rcs, 'attention_aggregate_refine')
    prob_ref = A.output_conv_refine(cost_ref)
    return A.prob2depth_upsample(prob_ref, D, ds, di)


def test_example0_multiview_3_views(A):
    """example/0 with --view_num 3.  Three comparisons against the CPU oracle:
      (1) the hot path on IDENTICAL inputs: stages I + II on the oracle's FEM features (fp16 tensor-core path), and
          stages III + IV on the oracle's stage-II results -> inside the north-star 0.1 % of the depth range;
      (2) the 2-D feature extractor on the real images (fp32 both sides);
      (3) images -> final depth map end to end.  On these un-normalised 0..255 images the fp32 arithmetic of the
          reference algorithm itself is only reproducible to ~5e-4 of the range after stage I (oracle in fp32 vs the same
          oracle in fp64: tests/test_gpu_cfg1.py::test_example2_twoview measures it) and the refinement stage's
          nearest-neighbour warps / visual-hull thresholds amplify that, so (3) is bounded by a multiple of that floor,
          not by 0.1 %."""
    from oracle import schedule as osch
    D = 128
    images, cams, _ = A.pipeline.load_example(os.path.join(EX, '0'), view_num=3)
    assert images.shape == (1, 3, 640, 960, 3) and cams.shape == (1, 3, 2, 4, 4)
    w = all_weights(A)
    A.variables.load_weights(w)
    assert A.FLAGS.precision == 'fp16'
    ref = osch.run_multiview(images, cams, D, w)
    # (1) identical inputs
    out = A.pipeline.run_multiview(cu(ref['features']), cu(cams), D, siamese=True, upsample=True)
    m2 = mae_over_range(npy(out['depth']), ref['depth_agg_init'], cams, D)
    mv = [mae_over_range(npy(a), b, cams, D) for a, b in zip(out['depth_views'], ref['depth_views'])]
    est, est_up = _final_from_stage2(A, ref, images, cams, D)
    m4 = mae_over_range(npy(est_up), ref['depth_refined_up'], cams, D)
    print("example/0 (3 views), identical inputs: stage II depth MAE / range %.3e, depth_views %s, stages III+IV %.3e"
          % (m2, ["%.3e" % v for v in mv], m4))
    # stages I + II (the fp16 tensor-core path): the north-star 0.1 %.  Stages III + IV run on the fp32 CUDA-core path; with
    # these weights the logits reach +-140 on real images, so fp32 rounding alone moves the refined soft-argmin by ~1e-3
    # of the range (the fp32 oracle against its own fp64 evaluation: 1.3e-3 on example/2, test_example2_twoview)
    assert m2 < 1e-3 and max(mv) < 1e-3, (m2, mv)
    assert m4 < 3e-3, m4
    # (2) FEM
    feats = A.fem.extract_features(cu(images))
    fe = float(np.abs(npy(feats) - ref['features']).max() / np.abs(ref['features']).max())
    print("example/0: FEM features max rel err %.3e" % fe)
    assert fe < 5e-3
    # (3) end to end from the images
    e2e = A.pipeline.run_example(cu(images), cu(cams), D)
    torch.cuda.synchronize()
    e2 = mae_over_range(npy(e2e['depth']), ref['depth_agg_init'], cams, D)
    e4 = mae_over_range(npy(e2e['depth_refined_up']), ref['depth_refined_up'], cams, D)
    p = torch.softmax(-torch.from_numpy(ref['refined_prob_volume_agg']), dim=1).max(dim=1).values.mean().item()
    print("example/0 end to end from images: stage II %.3e, final (stage IV, x4) %.3e of the range, mean peak probability %.3f"
          % (e2, e4, p))
    assert tuple(e2e['depth_refined_up'].shape) == (1, 640, 960, 1)
    assert e2 < 2e-3 and e4 < 1.5e-2, (e2, e4)
    # host epilogue (example.py:183-186): depth = 1 / inverse depth, inside the swept range
    pred = npy(e2e['pred'])
    lo, hi = float(cams[0, 0, 1, 3, 0]), float(cams[0, 0, 1, 3, 0] + (D - 1) * cams[0, 0, 1, 3, 1])
    assert np.isfinite(pred).all() and pred.min() >= 1.0 / hi - 1e-4 and pred.max() <= 1.0 / lo + 1e-4


def test_example2_twoview(A):
    """example/2 (only two views exist -> the two-view network TVSNet, example.py:344-347).  End to end from the images
    against the fp32 oracle AND against the same oracle evaluated in fp64: the CUDA path must sit at the fp32
    reproducibility floor of the reference algorithm on this input (its error against the fp64 result is at most 3x the
    fp32 oracle's own), and the refined depth map stays within 0.5 % of the range of the fp32 oracle."""
    import oracle.fem as ofem
    import oracle.homography_warping as ohw
    import oracle.model as om
    import oracle.network as onet
    import oracle.refine as oref
    from oracle import schedule as osch
    D = 128
    images, cams, _ = A.pipeline.load_example(os.path.join(EX, '2'), view_num=3)     # only 2 views exist -> two-view path
    assert images.shape == (1, 2, 480, 640, 3)
    w = all_weights(A)
    A.variables.load_weights(w)
    out = A.pipeline.run_example(cu(images), cu(cams), D)
    torch.cuda.synchronize()
    ref = osch.run_twoview(images, cams, D, w)
    mods = (ofem, ohw, om, onet, oref, osch)
    try:
        for m in mods:
            m.F32 = np.float64                      # fp64 twin of the oracle (same code, double arithmetic)
        ref64 = osch.run_twoview(images.astype(np.float64), cams.astype(np.float64), D, w)
    finally:
        for m in mods:
            m.F32 = np.float32
    m_b2 = mae_over_range(npy(A.prob2depth(out['refined_prob_volume'], D, cu(cams[:, 0, 1, 3, 0]), cu(cams[:, 0, 1, 3, 1]))),
                          ref['depth_refined'], cams, D)
    m = mae_over_range(npy(out['depth_refined_up']), ref['depth_refined_up'], cams, D)
    ours64 = mae_over_range(npy(out['depth_refined_up']), ref64['depth_refined_up'], cams, D)
    floor = mae_over_range(ref['depth_refined_up'], ref64['depth_refined_up'], cams, D)
    print("example/2 two-view: refined depth MAE / range vs fp32 oracle %.3e (low-res %.3e); vs fp64 oracle: ours %.3e, "
          "fp32 oracle %.3e" % (m, m_b2, ours64, floor))
    assert tuple(out['depth_refined_up'].shape) == (1, 480, 640, 1)
    assert ours64 < 3 * floor + 2e-4, (ours64, floor)
    assert m < 5e-3, m
