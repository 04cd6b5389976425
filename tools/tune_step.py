"""Coordinate descent over the launch-shape knobs (env variables read by libatvs.so at launch time) on the WHOLE cfg2
step (CUDA-graph replay, 4 streams): what is best for a kernel alone is not what is best when the passes of a step share
the SMs.  Prints every evaluation and the best setting found.

    python tools/tune_step.py [rounds]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import atvsnet_b200 as A

# round 2, balanced plane ranges (ring_common.cuh RingSpan): CTAs per launch instead of z-segment lengths
KNOBS = [
    ("ATVS_DRING_CTAS", [None, 4, 8, 12, 16, 24]),
    ("ATVS_S2_CTAS_8_16", [None, 12, 16, 24, 32, 48]),
    ("ATVS_RING_CTAS_16_16", [None, 6, 8, 12, 16, 24]),
    ("ATVS_RING_CTAS_8_8", [None, 32, 48, 64, 80, 100, 148]),
    ("ATVS_RING_CTAS_32_8", [None, 64, 100, 148, 200]),
    ("ATVS_RING_CTAS_8_16", [None, 64, 100, 148, 296]),
    ("ATVS_RING_CTAS_8_1", [None, 64, 100, 148, 200]),
    ("ATVS_RING_MINPLANES", [28, 40, 56]),
]
START = {"ATVS_RING_MINPLANES": 40, "ATVS_RING_CTAS_8_8": 100, "ATVS_RING_CTAS_8_1": 200, "ATVS_DRING_CTAS": 16}

nv, h, w, D = 5, 128, 160, 128
A.variables.load_weights(A.variables.synthetic_weights())
cams = torch.from_numpy(A.synthetic.orbit_cams(nv, h, w, D)[None]).cuda()
feats = torch.from_numpy(A.synthetic.smooth_features(nv, h, w, 32, seed=0)[None]).cuda()


def apply(cfg):
    for k, _ in KNOBS:
        os.environ.pop(k, None)
    for k, v in cfg.items():
        if v is not None:
            os.environ[k] = str(v)
    A.pipeline.CONCURRENT_PASSES = int(cfg.get("ATVS_PASSES") or 8)


def measure(cfg, steps=12):
    apply(cfg)
    step = lambda: A.pipeline.run_multiview(feats, cams, D, siamese=True)['depth_up']
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            step()
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    del g
    return best


cfg = dict(START)
best = measure(cfg)
print("start", json.dumps(cfg), "%.3f ms" % best, flush=True)
print("defaults", "%.3f ms" % measure({}), flush=True)
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for r in range(rounds):
    for name, values in KNOBS:
        for v in values:
            if cfg.get(name) == v:
                continue
            trial = dict(cfg)
            trial[name] = v
            t = measure(trial)
            print("  %s=%s -> %.3f ms" % (name, v, t), flush=True)
            if t < best - 0.01:
                best, cfg = t, trial
        print("round %d after %s: best %.3f ms %s" % (r, name, best, json.dumps(cfg)), flush=True)
print("BEST", "%.3f ms" % best, json.dumps(cfg))
