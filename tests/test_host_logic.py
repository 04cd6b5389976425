"""CPU: host-side logic of the product package that needs no device - variable naming,
synthetic generators, view sharding, graph recording / fusion plan - and the world_size-2
gloo check of the sharded attention combine."""
import os
import sys

import numpy as np
import pytest
import torch

import atvsnet_b200 as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_variable_names_match_checkpoint_convention(gweights):
    w = A.variables.synthetic_weights(seed=1)
    conv_like = {k: v.shape for k, v in w.items() if 'batch_normalization' not in k}
    assert set(conv_like) == set(gweights)
    for k in gweights:
        assert conv_like[k] == gweights[k].shape, k
    assert w['conv_b0_1_0/conv3d/kernel'].shape == (3, 3, 3, 64, 16)
    assert w['conv_b1_4_0/conv3d_transpose/kernel'].shape == (3, 3, 3, 32, 64)       # [.., Cout, Cin]
    assert w['conv_b2_6_2/kernel'].shape == (3, 3, 3, 8, 1)
    assert w['attention_aggregate/attention_activation/weight_shared'].shape == (3, 3, 3, 8, 8)
    n_params = sum(v.size for k, v in w.items() if k.startswith('conv_b') and 'batch_normalization' not in k)
    assert n_params == 912600
    assert 'conv_b0_1_0/batch_normalization/moving_mean' in w                          # restored but never read (F4)


def test_fem_variable_names_equal_the_reference_graph_list():
    import json
    rec = {k: tuple(v) for k, v in json.load(open(os.path.join(ROOT, 'tests', 'golden', 'fem_variables.json'))).items()}
    assert A.variables.fem_variable_shapes() == rec
    w = A.variables.synthetic_fem_weights()
    assert set(w) == set(rec) and all(w[k].shape == rec[k] and w[k].dtype == np.float32 for k in rec)


def test_refine_variable_names_equal_the_reference_code_list():
    import json
    rec = {k: tuple(v) for k, v in json.load(open(os.path.join(ROOT, 'tests', 'golden', 'refine_variables.json'))).items()}
    assert A.variables.refine_variable_shapes() == rec
    w = A.variables.synthetic_refine_weights()
    assert set(w) == set(rec) and all(w[k].shape == rec[k] for k in rec)


def test_flags_defaults_follow_reference():
    assert A.FLAGS.max_d == 128 and A.FLAGS.view_num == 5 and A.FLAGS.batch_size == 1
    assert A.FLAGS.inverse_depth is True and A.FLAGS.sample_scale == 0.25


def test_orbit_rig_and_features():
    cams = A.synthetic.orbit_cams(5, 128, 160, 128)
    assert cams.shape == (5, 2, 4, 4) and cams.dtype == np.float32
    assert np.allclose(cams[0, 0], np.eye(4))
    for i in range(5):
        R = cams[i, 0, :3, :3].astype(np.float64)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-6)
        # every camera looks at the pivot: it projects to the principal point
        P = R @ np.array([0, 0, 5.0]) + cams[i, 0, :3, 3]
        assert np.allclose(P, [0, 0, 5.0], atol=1e-5)
    assert np.isclose(cams[0, 1, 3, 0], 0.05) and np.isclose(cams[0, 1, 3, 1], 0.4224 / 128)
    f = A.synthetic.smooth_features(2, 16, 24, 32, seed=0)
    assert f.shape == (2, 16, 24, 32) and abs(f.mean()) < 1e-3 and abs(f.std() - 1) < 1e-3
    assert np.array_equal(f, A.synthetic.smooth_features(2, 16, 24, 32, seed=0))


def test_shard_views():
    assert A.pipeline.shard_views(9, 0, 4) == [1, 5]
    assert A.pipeline.shard_views(9, 3, 4) == [4, 8]
    got = sorted(v for r in range(8) for v in A.pipeline.shard_views(9, r, 8))
    assert got == list(range(1, 9))
    assert A.pipeline.shard_views(5, 0, 1) == [1, 2, 3, 4]


def test_graph_recording_matches_reference_layer_names():
    t = A.StackedUNet_prob({'data': torch.zeros(1, 8, 8, 8, 64)})      # recording only: nothing runs yet
    names = list(t.nodes)
    convs = [n for n in names if t.nodes[n].kind in ('conv_bn', 'deconv_bn', 'conv')]
    assert len(convs) == 31
    assert names[-1] == 'conv_b2_6_2' and t.nodes['conv_b2_6_2'].inputs == ['conv_b2_6_1']
    assert t.nodes['conv_b1_4_1'].inputs == ['conv_b1_4_0', 'conv_b1_2_1', 'conv_b0_2_1']
    assert t.nodes['conv_b2_5_1'].inputs == ['conv_b2_5_0', 'conv_b2_1_1', 'conv_b0_1_1']
    assert t.nodes['conv_b1_1_1_concat'].inputs == ['conv_b1_1_0', 'conv_b0_5_0']
    assert t.nodes['conv_b0_4_0'].params == dict(filters=32, stride=2, relu=True)
    with pytest.raises(KeyError):
        t.feed('no_such_layer')
    u = A.StackedUNet({'data': torch.zeros(1, 8, 8, 8, 64)})
    assert list(u.nodes)[-1] == 'conv_b2_6_1'


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
    from gen_common import golden_weights
    from oracle import network as onet
    w = golden_weights(7)
    rng = np.random.default_rng(42)
    xs = rng.standard_normal((1, 4, 6, 6, 8, 5)).astype(np.float32)          # 5 source views
    mine = [v - 1 for v in A.pipeline.shard_views(6, rank, world)]
    wu = w['attention_aggregate/attention_activation/weight_unique']
    ws = w['attention_aggregate/attention_activation/weight_shared']
    # local logits l_n = relu(conv_u x_n) - relu(conv_s x_n): the +sum_m s_m term cancels in the softmax
    l = [torch.from_numpy(onet.relu(onet.conv3d(xs[..., n], wu)) - onet.relu(onet.conv3d(xs[..., n], ws)))
         for n in mine]
    lmax = torch.stack(l).max(dim=0).values
    dist.all_reduce(lmax, op=dist.ReduceOp.MAX)
    e = [torch.exp(v - lmax) for v in l]
    num = sum(ei * torch.from_numpy(xs[..., n]) for ei, n in zip(e, mine))
    den = sum(e)
    nd = torch.cat([num, den], dim=-1)
    # the product's exchange step (pipeline.reduce_partials): reduce-scatter (all-reduce + slice on gloo), finish on the
    # owned slab, all-gather of the result; 4*6*6 = 144 voxels and a ragged 143-voxel case (padding path)
    flat = nd.reshape(-1, 16)
    fin = lambda s_: s_[:, :8] / s_[:, 8:]
    out = A.pipeline.reduce_partials(flat.clone(), world, rank, None, fin, torch.float32).reshape(1, 4, 6, 6, 8).numpy()
    ragged = A.pipeline.reduce_partials(flat[:143].clone(), world, rank, None, fin, torch.float32).numpy()
    assert np.array_equal(ragged, out.reshape(-1, 8)[:143])
    # a bf16-rounded shift that is the same on every rank is as good as the exact max
    l16 = torch.stack(l).max(dim=0).values.to(torch.bfloat16)
    dist.all_reduce(l16, op=dist.ReduceOp.MAX)
    e2 = [torch.exp(v - l16.float()) for v in l]
    nd2 = torch.cat([sum(ei * torch.from_numpy(xs[..., n]) for ei, n in zip(e2, mine)), sum(e2)], dim=-1).reshape(-1, 16)
    out2 = A.pipeline.reduce_partials(nd2, world, rank, None, fin, torch.float32).reshape(1, 4, 6, 6, 8).numpy()
    assert np.abs(out2 - out).max() < 1e-5 * np.abs(out).max()
    if rank == 0:
        ref = onet.AttAggregation_keepchannel({'data': xs}, w).get_output()
        q.put(float(np.abs(out - ref).max() / np.abs(ref).max()))
    dist.destroy_process_group()


def test_sharded_attention_combine_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert err < 1e-5, err


def test_attention_fused_applicability():
    """network.attention_fused_ok: the one-kernel AAM covers 2..8 views of (B,D,H,W,8) 16-bit volumes with D >= 3 and
    H, W >= 8; everything else (fp32, other channel counts, one view, tiny planes, the flag off) takes the two-kernel path."""
    import torch
    import atvsnet_b200 as A
    N = A.network
    mk = lambda shape, dt=torch.float16: torch.empty(shape, dtype=dt, device='meta')
    ok = mk((1, 16, 32, 40, 8))
    assert N.attention_fused_ok([ok] * 2) and N.attention_fused_ok([ok] * 4) and N.attention_fused_ok([ok] * 8)
    assert N.attention_fused_ok([mk((2, 3, 8, 8, 8), torch.bfloat16)] * 3)
    assert not N.attention_fused_ok([ok])                                   # softmax over one view: identity path
    assert not N.attention_fused_ok([ok] * 9)
    assert not N.attention_fused_ok([mk((1, 16, 32, 40, 8), torch.float32)] * 4)
    assert not N.attention_fused_ok([mk((1, 16, 32, 40, 16))] * 4)
    assert not N.attention_fused_ok([mk((1, 2, 32, 40, 8))] * 4)
    assert not N.attention_fused_ok([mk((1, 16, 4, 40, 8))] * 4)
    A.FLAGS.attention_fused = False
    try:
        assert not N.attention_fused_ok([ok] * 4)
    finally:
        A.FLAGS.attention_fused = True


def test_shallow_cache_computes_every_view_once(monkeypatch):
    """refine.shallow_cache: inside the context the shallow features of a view are computed once however many
    (reference, source) pairs ask for them; outside it every call computes both (the reference's behaviour)."""
    import torch
    import atvsnet_b200 as A
    R = A.refine
    calls = []

    def fake(image):
        calls.append(image.data_ptr())
        return image.sum()
    monkeypatch.setattr(R, 'shallow_features', fake)
    images = torch.arange(5 * 4, dtype=torch.float32).reshape(1, 5, 2, 2, 1)
    for v in range(1, 5):
        R.extract_feature_shallow(images, 0, v)
    assert len(calls) == 8
    del calls[:]
    with R.shallow_cache():
        outs = [R.extract_feature_shallow(images, 0, v) for v in range(1, 5)]
    assert len(calls) == 5                                                  # view 0 once + 4 sources
    assert all(float(o[0]) == float(images[:, 0].sum()) for o in outs)
    assert [float(o[1]) for o in outs] == [float(images[:, v].sum()) for v in range(1, 5)]
    assert R._SHALLOW_MEMO is None                                          # the context cleans up after itself
