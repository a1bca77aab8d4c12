"""[phase traces need a library built with `make -C py4cast_b200/csrc EXTRA=-DP4C_PHASE_TRACE`]
Phase timing (SM clocks) of ONE small forward kernel launch: where do ~15 us go?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from py4cast_b200 import _lib, ops
from py4cast_b200.graph_build import build_mesh_graph, unit_meshgrid
from py4cast_b200.nlam import InteractionNet

dev = torch.device("cuda:0")
lib = _lib.load()
graph = build_mesh_graph(unit_meshgrid(512, 640), True, 3)
names = ["start", "setup done", "row_info", "gather issued", "X ready", "MMA1 done", "ep1 done", "MMA2 done", "ep2 done", "exit"]
for label, es in (("same[2] (544 e, 81 n)", graph.m2m[2]), ("same[1] (5512 e, 729 n)", graph.m2m[1]), ("up[1] 729 e", graph.mesh_up[1])):
    net = InteractionNet(es, 64).to(dev)
    packer = ops.WeightPacker([m.meta for m in net.modules() if hasattr(m, "meta")])
    packer.pack()
    x = torch.randn(1, es.num_send, 64, device=dev).bfloat16()
    r = torch.randn(1, es.num_rec, 64, device=dev).bfloat16()
    e = torch.randn(1, es.num_edges, 64, device=dev).bfloat16()
    trace = torch.zeros(16, dtype=torch.int64, device=dev)
    with torch.no_grad():
        for _ in range(3):
            net(x, r, e)
        torch.cuda.synchronize()
        lib.p4c_set_trace_buffer(trace.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        net(x, r, e)
        e1.record()
        torch.cuda.synchronize()
        lib.p4c_set_trace_buffer(None)
    t = trace.cpu().tolist()
    print(f"{label}: 2 launches {e0.elapsed_time(e1)*1e3:.1f} us; LAST kernel (node MLP, seg prologue) phases in cycles:")
    print("   " + ", ".join(f"{n}:+{t[i]-t[0]}" for i, n in enumerate(names)))
    print("   seg_sum_tile: " + ", ".join(f"{n}:+{t[i]-t[0]}" for i, n in zip(range(10, 15), ["entry", "setup", "loop done", "syncwarp", "stored"])) + f", iterations {t[15]}")
