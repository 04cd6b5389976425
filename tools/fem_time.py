"""FEM (ResNetDS2SPP, fp32 parity path) on the 5 views of a cfg2 frame: ms per frame, graph replay."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import atvsnet_b200 as A
A.variables.load_weights(A.variables.synthetic_fem_weights())
img = torch.rand(1, 5, 512, 640, 3, device='cuda') * 255
for bv in ((True,) if os.environ.get("BATCH_ONLY") else (True, False)):
    A.fem.BATCH_VIEWS = bv
    for _ in range(2):
        A.fem.extract_features(img)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); st = torch.cuda.Stream(); st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            o = A.fem.extract_features(img)
    torch.cuda.current_stream().wait_stream(st)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"batch_views": bv, "ms_per_frame": e0.elapsed_time(e1) / 5, "tflops_fp32": 5 * 84.0 / (e0.elapsed_time(e1) / 5)}), flush=True)
    del g, o
