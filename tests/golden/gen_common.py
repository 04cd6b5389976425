"""Seeded inputs shared by make_golden.py (generation, build container only) and the tests
(which only need ``golden_weights`` - the weights are too large to store as fixtures)."""
import numpy as np


def crm_layers():
    t = []
    for b in range(3):
        p = 'conv_b%d' % b
        c0 = 64 if b == 0 else 8
        t += [(p + '_1_0', 'c', c0, 16), (p + '_2_0', 'c', 16, 32), (p + '_3_0', 'c', 32, 64),
              (p + '_0_1', 'c', c0, 8), (p + '_1_1', 'c', 16, 16), (p + '_2_1', 'c', 32, 32),
              (p + '_3_1', 'c', 64, 64), (p + '_4_0', 'd', 64, 32), (p + '_5_0', 'd', 32, 16),
              (p + '_6_0', 'd', 16, 8)]
    return t


def golden_weights(seed=7):
    """name -> fp32 array with the checkpoint's names and TF layouts."""
    rng = np.random.default_rng(seed)
    w = {}
    for name, kind, cin, cout in crm_layers():
        if kind == 'c':
            w[name + '/conv3d/kernel'] = (rng.standard_normal((3, 3, 3, cin, cout)) * np.sqrt(2.0 / (27 * cin))).astype(np.float32)
        else:
            w[name + '/conv3d_transpose/kernel'] = (rng.standard_normal((3, 3, 3, cout, cin)) * np.sqrt(16.0 / (27 * cin))).astype(np.float32)
    w['conv_b2_6_2/kernel'] = (rng.standard_normal((3, 3, 3, 8, 1)) * 0.3).astype(np.float32)
    for scope, outc in (('attention_aggregate', 'attention_prob_vol'),
                        ('attention_aggregate_refine', 'attention_prob_vol_refine')):
        for nm in ('weight_unique', 'weight_shared'):
            w[scope + '/attention_activation/' + nm] = (rng.standard_normal((3, 3, 3, 8, 8)) * 0.1).astype(np.float32)
        w[outc + '/kernel'] = (rng.standard_normal((3, 3, 3, 8, 1)) * 0.3).astype(np.float32)
    return w


def smooth(rng, shape, passes=2):
    x = rng.standard_normal(shape).astype(np.float32)
    for _ in range(passes):
        for ax in range(1, len(shape) - 1):
            x = (np.roll(x, 1, ax) + x + np.roll(x, -1, ax)) / np.float32(3)
    return (x / x.std()).astype(np.float32)


def fem_variable_shapes():
    """name -> shape of every variable the reference's ResNetDS2SPP graph asks for (written by make_golden_fem.py
    from the reference code itself)."""
    import json
    import os
    return {k: tuple(v) for k, v in json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                                 'fem_variables.json'))).items()}


def fem_weights(seed=7):
    """seeded FEM weights under the checkpoint's names: He-normal kernels, small biases / betas."""
    w = {}
    for i, (name, shape) in enumerate(sorted(fem_variable_shapes().items())):
        rng = np.random.default_rng([seed, i])
        if len(shape) == 4:
            w[name] = (rng.standard_normal(shape) * np.sqrt(2.0 / np.prod(shape[:-1]))).astype(np.float32)
        else:
            w[name] = (rng.standard_normal(shape) * 0.1).astype(np.float32)
    return w


def named_weights(json_name, seed):
    """seeded weights for a recorded variable list (name -> shape JSON next to this file): He-normal kernels (4-D and
    5-D; transposed-conv kernels [..,Cout,Cin] use fan-in of their LAST axis), small biases / betas."""
    import json
    import os
    shapes = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), json_name)))
    w = {}
    for i, (name, shape) in enumerate(sorted(shapes.items())):
        rng = np.random.default_rng([seed, i])
        shape = tuple(shape)
        if len(shape) >= 4:
            fan_in = np.prod(shape[:-2]) * (shape[-1] if 'transpose' in name else shape[-2])
            w[name] = (rng.standard_normal(shape) * np.sqrt(2.0 / fan_in)).astype(np.float32)
        else:
            w[name] = (rng.standard_normal(shape) * 0.1).astype(np.float32)
    return w
