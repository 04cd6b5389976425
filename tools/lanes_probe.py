"""Throughput of the cfg2 step with L captured graphs (separate buffers) replayed on L streams, frames dealt round-robin:
does a second frame in flight fill the tails of the first (stage II is a serial chain, stage I ramps up on K1)?
    python tools/lanes_probe.py [lanes ...]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import atvsnet_b200 as A
import bench

dev = torch.device('cuda:0')
A.variables.load_weights(A.variables.synthetic_weights(), device=dev)
feats, cams, D = bench.make_inputs('cfg2', frame_seed=0)
feats_d = torch.from_numpy(feats).to(dev)
cams_d = torch.from_numpy(cams).to(dev)
for lanes in [int(a) for a in sys.argv[1:]] or [1, 2, 3]:
    for passes in (8, 4):
        A.pipeline.CONCURRENT_PASSES = passes
        graphs, streams, outs = [], [], []
        for _ in range(3):
            A.pipeline.run_multiview(feats_d, cams_d, D, siamese=True)
        torch.cuda.synchronize()
        for l in range(lanes):
            s = torch.cuda.Stream(device=dev)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    o = A.pipeline.run_multiview(feats_d, cams_d, D, siamese=True)['depth_up']
            torch.cuda.current_stream().wait_stream(s)
            graphs.append(g); streams.append(s); outs.append(o)
        torch.cuda.synchronize()
        nfr = 30
        def run():
            for i in range(nfr):
                with torch.cuda.stream(streams[i % lanes]):
                    graphs[i % lanes].replay()
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams: s.wait_event(e0)
        run()
        for s in streams: torch.cuda.current_stream().wait_stream(s)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / nfr
        same = all(torch.equal(outs[0], o) for o in outs)
        print(json.dumps({"lanes": lanes, "passes": passes, "ms_per_map": ms, "maps_s": 1000 / ms, "same": same}), flush=True)
        del graphs, outs
        torch.cuda.synchronize(); torch.cuda.empty_cache()
