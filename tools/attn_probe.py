"""K2 at cfg2 size (4 views, 128x128x160x8): the one-kernel AAM (atvs_attention_fused) against the two-kernel path
(8 -> 16 convolution per view + atvs_attention_raw), graph replay of 10, CUDA events.
    python tools/attn_probe.py [nviews D H W]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import atvsnet_b200 as A

nv, D, H, W = [int(a) for a in sys.argv[1:5]] if len(sys.argv) >= 5 else (4, 128, 128, 160)
dev = torch.device('cuda:0')
A.variables.load_weights(A.variables.synthetic_weights(), device=dev)
views = [torch.randn(1, D, H, W, 8, device=dev).clamp_(min=-0.5).to(torch.float16) for _ in range(nv)]


def timed(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(iters):
                fn()
    torch.cuda.current_stream().wait_stream(st)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


V = D * H * W
alg = nv * V * 16 + V * 32
for fused in ((True,) if os.environ.get("FUSED_ONLY") else (True, False)):
    A.FLAGS.attention_fused = fused
    us = timed(lambda: A.cost_volume_aggregation(views, keepchannel=True))
    print(json.dumps({"fused": fused, "us": us, "knobs": {k: v for k, v in os.environ.items() if k.startswith('ATVS_ATTN')},
                      "algorithmic_GBs": alg / us / 1e3}), flush=True)
