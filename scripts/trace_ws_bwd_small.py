"""[phase traces need a library built with `make -C py4cast_b200/csrc EXTRA=-DP4C_PHASE_TRACE`]
Kernel-level timeline (SM clocks of CTA 0) of the pipelined backward on the MID-SIZE calls of the mesh
levels (level-0 InteractionNet at the bench batch): where do the ~70 us of an 820-tile node MLP go?
P4C_WB_TRACE_NBLK=2 traces the node MLP (last launch with two K-blocks), =3 the edge MLP."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from py4cast_b200 import _lib, ops
from py4cast_b200.graph_build import build_mesh_graph, unit_meshgrid
from py4cast_b200.nlam import InteractionNet

dev = torch.device("cuda:0")
lib = _lib.load()
graph = build_mesh_graph(unit_meshgrid(512, 640), True, 3)
B = int(os.environ.get("B", "16"))
KN = ["entry", "setup done", "first tile top", "tiles done", "sync", "partials", "exit"]
CNAMES = ["top", "h1_full", "a1 out", "h2_full", "dh2 out", "colsums", "da1_full", "dh1 out", "dh1 stored", "dx_full", "end"]
for label, es, upd in (("same[0]", graph.m2m[0], True), ("up[1]", graph.mesh_up[1], False), ("m2g", graph.m2g, False)):
    torch.manual_seed(0)
    net = InteractionNet(es, 64, update_edges=upd).to(dev)
    ops.WeightPacker([m.meta for m in net.modules() if hasattr(m, "meta")]).pack()
    send = torch.randn(B, es.num_send, 64, device=dev).bfloat16().requires_grad_(True)
    rec = send if label.startswith("same") else torch.randn(B, es.num_rec, 64, device=dev).bfloat16().requires_grad_(True)
    edge = torch.randn(B, es.num_edges, 64, device=dev).bfloat16().requires_grad_(True)
    trace = torch.zeros(64, dtype=torch.int64, device=dev)
    for it in range(4):
        out = net(send, rec, edge)
        outs = [out[0], out[1]] if upd else [out]
        gs = [torch.ones_like(o) for o in outs]
        torch.cuda.synchronize()
        if it == 3:
            lib.p4c_set_trace_buffer(trace.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.autograd.backward(outs, gs)
        e1.record()
        torch.cuda.synchronize()
        lib.p4c_set_trace_buffer(None)
    t = trace.cpu().tolist()
    print(f"{label} batch {B}: rows send {es.num_send} rec {es.num_rec} edges {es.num_edges}; backward {e0.elapsed_time(e1)*1e3:.1f} us", flush=True)
    if t[48]:
        print(f"   traced launch: {t[55]} tiles in CTA 0, {t[57]-t[56]} ns entry -> exit (globaltimer)")
        print("   kernel : " + ", ".join(f"{n}:+{t[48+i]-t[48]}" for i, n in enumerate(KN)))
        print("   tile ends (group 0): " + ", ".join(f"+{v-t[48]}" for v in t[58:64] if v))
        if t[0]:
            print("   tile 2 : " + ", ".join(f"{n}:+{t[i]-t[48]}" for i, n in enumerate(CNAMES) if t[i]))
            extra = {13: "a: TMEM load done", 14: "c: TMEM load done", 11: "g: first TMEM load done", 15: "g: group sums done", 12: "g: slots done"}
            print("   tile 2 (inside the phases): " + ", ".join(f"{n}:+{t[i]-t[48]}" for i, n in extra.items() if t[i]))
