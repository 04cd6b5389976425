"""Premise check for batching the stage-I passes: the CRM on ONE batch of P cost volumes (one stream, BN statistics over
the batch - not the reference semantics, same work) against P single-sample passes on P streams (what the step does).
    python tools/batch_probe.py [P]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import atvsnet_b200 as A
from atvsnet_b200 import network as N, _lib as L
from atvsnet_b200.atvsnet import StackedUNet_prob

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
D, h, w = 128, 128, 160
dev = torch.device('cuda:0')
A.variables.load_weights(A.variables.synthetic_weights(), device=dev)
ref = torch.randn(P, h, w, 32, device=dev)
warped = torch.randn(P, D, h, w, 32, device=dev).to(torch.float16)


def crm(r, wv):
    t = StackedUNet_prob({'data': N.SplitCostVolume(r, wv)}, outputs=('conv_b2_6_1', 'conv_b2_6_2'))
    return t.get_output_by_name('conv_b2_6_1'), t.get_output()


def graph_time(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            out = fn()
    torch.cuda.current_stream().wait_stream(st)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def batched():
    A._lib.call("atvs_set_concurrency", 1)
    return crm(ref, warped)


streams = [torch.cuda.Stream() for _ in range(P)]


def streamed():
    A._lib.call("atvs_set_concurrency", P)
    main = torch.cuda.current_stream()
    outs = []
    for s in streams:
        s.wait_stream(main)
    for i, s in enumerate(streams):
        with torch.cuda.stream(s):
            outs.append(crm(ref[i:i + 1], warped[i:i + 1]))
    for s in streams:
        main.wait_stream(s)
    A._lib.call("atvs_set_concurrency", 1)
    return outs


crm(ref[:1], warped[:1]); torch.cuda.synchronize()     # packs weights on the main stream
if os.environ.get('BATCH_ONLY'):      # for a launch list under ncu: one eager batched pass
    batched(); torch.cuda.synchronize()
    batched(); torch.cuda.synchronize()
    sys.exit(0)
tb = graph_time(batched)
ts = graph_time(streamed)
print(json.dumps({"passes": P, "batched_ms": tb, "streamed_ms": ts}))
