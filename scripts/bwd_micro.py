"""[phase traces need a library built with `make -C py4cast_b200/csrc EXTRA=-DP4C_PHASE_TRACE`]
Backward of the m2g / g2m InteractionNets of the Titan-shaped HiLAM graph in isolation, classic
tensor-core engine against the pipelined recomputing one (CUDA events around the whole backward =
node MLP + edge MLP + segment-sum passes), and the phase trace of the pipelined edge kernel
(SM clocks of CTA 0, tile 2: compute group 0, MMA lane, producer warp 0)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CNAMES = ["top", "h1_full", "a1 out", "h2_full", "dh2 out", "colsums", "da1_full", "dh1 out", "dh1 stored",
          "dx_full", "end"]
MNAMES = ["G0", "G1", "G2/3", "G4", "G5", "x_full seen", "a1_full seen", "dh2_full seen", "dh1_full seen"]
PNAMES = ["top", "x_empty", "issued", "landed"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--engines", default="classic,pipelined")
    args = ap.parse_args()
    from py4cast_b200 import _lib, ops
    from py4cast_b200.graph_build import build_mesh_graph, unit_meshgrid
    from py4cast_b200.nlam import InteractionNet

    dev = torch.device("cuda:0")
    lib = _lib.load()
    graph = build_mesh_graph(unit_meshgrid(512, 640), True, 3)
    B = args.batch
    for label, es, upd in (("m2g", graph.m2g, False), ("g2m", graph.g2m, False), ("same[0]", graph.m2m[0], True)):
        if args.only and label != args.only:
            continue
        torch.manual_seed(0)
        net = InteractionNet(es, 64, update_edges=upd).to(dev)
        ops.WeightPacker([m.meta for m in net.modules() if hasattr(m, "meta")]).pack()
        g = torch.Generator(device="cuda").manual_seed(0)
        send = torch.randn(B, es.num_send, 64, device=dev, generator=g).bfloat16().requires_grad_(True)
        rec = send if label == "same[0]" else torch.randn(B, es.num_rec, 64, device=dev, generator=g).bfloat16().requires_grad_(True)
        if upd:
            edge = torch.randn(B, es.num_edges, 64, device=dev, generator=g).bfloat16().requires_grad_(True)
        else:
            edge = torch.randn(es.num_edges, 64, device=dev, generator=g).bfloat16().requires_grad_(True)
        for engine, thr in (("classic", 1 << 40), ("pipelined", 0)):
            if engine not in args.engines.split(","):
                continue
            lib.p4c_set_ws_bwd_min_tiles(thr)
            tf, tb = [], []
            trace = torch.zeros(64, dtype=torch.int64, device=dev)
            for it in range(args.iters + 2):
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                if it == args.iters + 1 and engine == "pipelined":
                    lib.p4c_set_trace_buffer(trace.data_ptr())
                e0.record()
                out = net(send, rec, edge)
                e1.record()
                if upd:
                    torch.autograd.backward([out[0], out[1]], [torch.ones_like(out[0]), torch.ones_like(out[1])])
                else:
                    out.backward(torch.ones_like(out))
                e2.record()
                torch.cuda.synchronize()
                lib.p4c_set_trace_buffer(None)
                if it >= min(2, args.iters + 1):
                    tf.append(e0.elapsed_time(e1))
                    tb.append(e1.elapsed_time(e2))
                del out
            print(f"{label} batch {B} {engine}: forward {min(tf):.3f} ms, backward {min(tb):.3f} ms", flush=True)
            if engine == "pipelined":
                t = trace.cpu().tolist()
                live = [v for v in t[:11] + t[16:25] + t[32:36] if v > 0]
                t0 = min(live) if live else 0
                print("   compute: " + ", ".join(f"{n}:+{t[i]-t0}" for i, n in enumerate(CNAMES) if t[i]))
                print("   mma    : " + ", ".join(f"{n}:+{t[16+i]-t0}" for i, n in enumerate(MNAMES) if t[16 + i]))
                print("   produce: " + ", ".join(f"{n}:+{t[32+i]-t0}" for i, n in enumerate(PNAMES) if t[32 + i]))
        lib.p4c_set_ws_bwd_min_tiles(0)


if __name__ == "__main__":
    main()
