"""Runs the m2g (or g2m) InteractionNet of the Titan-shaped HiLAM graph in isolation:
edge-MLP fwd, segment sum, node-MLP fwd, and the backward kernels.  Used under ncu, and
stand-alone it prints CUDA-event timings per phase."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="m2g", choices=["m2g", "g2m"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--batch", type=int, default=1)
    args = ap.parse_args()
    from py4cast_b200.graph_build import build_mesh_graph, unit_meshgrid
    from py4cast_b200.nlam import InteractionNet

    dev = torch.device("cuda:0")
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    graph = build_mesh_graph(unit_meshgrid(512, 640), True, 3)
    es = graph.m2g if args.which == "m2g" else graph.g2m
    torch.manual_seed(0)
    net = InteractionNet(es, 64, update_edges=False).to(dev)
    from py4cast_b200 import ops
    if dt == torch.bfloat16:  # weight images, as in a rollout (the pipelined forward needs them)
        ops.WeightPacker([m.meta for m in net.modules() if hasattr(m, "meta")]).pack()
    B = args.batch
    g = torch.Generator(device="cuda").manual_seed(0)
    send = torch.randn(B, es.num_send, 64, device=dev, generator=g).to(dt).requires_grad_(True)
    rec = torch.randn(B, es.num_rec, 64, device=dev, generator=g).to(dt).requires_grad_(True)
    edge = torch.randn(es.num_edges, 64, device=dev, generator=g).to(dt).requires_grad_(True)
    gout = torch.randn(B, es.num_rec, 64, device=dev, generator=g).to(dt)
    for it in range(args.iters):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        out = net(send, rec, edge)
        e1.record()
        out.backward(gout)
        e2.record()
        torch.cuda.synchronize()
        print(f"iter {it}: forward {e0.elapsed_time(e1):.3f} ms, backward {e1.elapsed_time(e2):.3f} ms")


if __name__ == "__main__":
    main()
