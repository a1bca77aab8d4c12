"""[phase traces need a library built with `make -C py4cast_b200/csrc EXTRA=-DP4C_PHASE_TRACE`]
Phase timing (SM clocks) of the tensor-core BACKWARD kernel: CTA 0, thread 0, tiles 0 and 2 of
the LAST backward-kernel launch (the edge MLP of an InteractionNet, or a plain MLP)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from py4cast_b200 import _lib, ops
from py4cast_b200.graph_build import build_mesh_graph, unit_meshgrid
from py4cast_b200.nlam import InteractionNet

dev = torch.device("cuda:0")
lib = _lib.load()
graph = build_mesh_graph(unit_meshgrid(512, 640), True, 3)
names = ["start", "setup", "pdl+rowinfo", "top(h1 in)", "loads used", "G1 done", "LN done", "G2/3 done", "dh1 done",
         "X in+sync", "G4/5 done", "tile end", "partials", "exit"]
B = int(os.environ.get("B", "2"))
for label, es, upd in (("m2g edge (1.3M e)", graph.m2g, False), ("same[0] (51520 e)", graph.m2m[0], True),
                       ("same[2] (544 e, 81 n)", graph.m2m[2], True), ("same[1] (5512 e, 729 n)", graph.m2m[1], True),
                       ("up[1] 729 e", graph.mesh_up[1], False)):
    net = InteractionNet(es, 64, update_edges=upd).to(dev)
    packer = ops.WeightPacker([m.meta for m in net.modules() if hasattr(m, "meta")])
    packer.pack()
    x = torch.randn(B, es.num_send, 64, device=dev).bfloat16().requires_grad_(True)
    r = torch.randn(B, es.num_rec, 64, device=dev).bfloat16().requires_grad_(True)
    e = torch.randn(B, es.num_edges, 64, device=dev).bfloat16().requires_grad_(True)
    trace = torch.zeros(64, dtype=torch.int64, device=dev)
    for i in range(3):
        out = net(x, r, e)
        loss_in = out if not upd else out[0]
        gout = torch.randn_like(loss_in)
        torch.cuda.synchronize()
        if i == 2:
            lib.p4c_set_trace_buffer(trace.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if upd:
            torch.autograd.backward([out[0], out[1]], [gout, torch.randn_like(out[1])])
        else:
            loss_in.backward(gout)
        e1.record()
        torch.cuda.synchronize()
    lib.p4c_set_trace_buffer(None)
    t = trace.cpu().tolist()
    print(f"{label}: backward (4 launches) {e0.elapsed_time(e1)*1e3:.1f} us; edge-MLP bwd kernel phases in cycles:")
    for base, nm in ((16, "tile 0"), (32, "tile 2")):
        if t[base + 3] == 0:
            continue
        t0 = t[16]
        print(f"   {nm}: " + ", ".join(f"{n}:+{t[base+i]-t0}" for i, n in enumerate(names) if t[base + i]))
