"""Runs the pipelined forward + backward of a dense MLP repeatedly on the same inputs and reports which outputs
are not bit-identical between runs (developer aid for race hunting)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from py4cast_b200 import _lib, ops
from test_gpu_ws_forward import _mlp, _weights

cuda = torch.device("cuda:0")
lib = _lib.load()
k1 = int(os.environ.get("K1", "128"))
rows = int(os.environ.get("ROWS", "30000"))
ref = _mlp(k1, seed=1)
ws = _weights(ref, True, cuda)
meta = ops.MlpMeta(ws, (k1,), 64)
ops.WeightPacker([meta]).pack()
x = torch.randn(2, rows, k1, device=cuda).bfloat16()
gy = torch.randn(2, rows, 64, device=cuda).bfloat16()
lib.p4c_set_ws_bwd_min_tiles(1)
names = ["y", "dx", "dW1", "db1", "dW2", "db2", "dgamma", "dbeta"]

def run():
    for w in ws:
        w.grad = None
    xx = x.clone().requires_grad_(True)
    y = ops.fused_mlp(xx, ws, meta=meta)
    y.backward(gy)
    torch.cuda.synchronize()
    return [y.detach().clone(), xx.grad.clone()] + [w.grad.clone() for w in ws]

base = run()
bad = {}
for it in range(int(os.environ.get("ITERS", "40"))):
    cur = run()
    for n, a, b in zip(names, base, cur):
        if not torch.equal(a, b):
            d = (a.float() - b.float()).abs()
            bad.setdefault(n, []).append((it, int((d > 0).sum()), float(d.max())))
print("mismatches:", {k: v[:4] for k, v in bad.items()} if bad else "none")
