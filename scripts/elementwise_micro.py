"""HBM-bound element-wise kernels of the path (AR step, loss) in isolation at the bench shapes:
CUDA-event time per call and algorithmic GB/s (SURVEY.md 8(d) byte counts) against the measured
HBM peak.  usage: elementwise_micro.py [--batch 8] [--steps 4]"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--grid", default="512x640")
    a = ap.parse_args()
    from py4cast_b200.datasets import synthetic_dataset_info
    from py4cast_b200.losses import ScaledLoss, WeightedLoss
    from py4cast_b200.namedtensor import NamedTensor
    from py4cast_b200.lightning import ar_concat

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pk = os.path.join(root, "MEASURED_PEAKS.json")
    peak = float(json.load(open(pk))["hbm_gbs"]) if os.path.isfile(pk) else 6650.0
    dev = torch.device("cuda:0")
    GH, GW = (int(v) for v in a.grid.split('x'))
    info = synthetic_dataset_info(GH, GW)
    B, T, N, F = a.batch, a.steps, GH * GW, info.weather_dim
    fn = info.shortnames["input_output"]
    names = ["batch", "timestep", "ngrid", "features"]
    pred = torch.randn(B, T, N, F, device=dev, requires_grad=True)
    tgt = torch.randn(B, T, N, F, device=dev)
    P, Tn = NamedTensor(pred, names, fn), NamedTensor(tgt, names, fn)

    class LM(torch.nn.Module):
        pass

    lm = LM().to(dev)
    interior = (1 - info.statics.border_mask).flatten(0, 1).to(dev)
    rows = []
    for label, loss in (("WeightedLoss(MSE)", WeightedLoss("MSELoss", reduction="none")),
                        ("ScaledLoss(L1)", ScaledLoss("L1Loss", reduction="none"))):
        loss.prepare(lm, interior, info)
        t_f = timeit(lambda: loss(P, Tn, None))
        out = loss(P, Tn, None)
        g = torch.ones_like(out)

        def fb():
            pred.grad = None
            loss(P, Tn, None).backward(g)

        t_fb = timeit(fb)
        by_f = 2 * 4 * B * T * N * F + 4 * N
        by_b = 3 * 4 * B * T * N * F
        rows.append((f"{label} forward", by_f, t_f))
        rows.append((f"{label} backward (fwd+bwd minus fwd)", by_b, max(t_fb - t_f, 1e-9)))
    # AR concat: x = cat(prev x T_in, statics, forcing) -> bf16 rows padded to 72
    prevs = [torch.randn(B, N, F, device=dev) for _ in range(2)]
    static = torch.randn(N, 4, device=dev)
    forcing = torch.randn(B, N, info.forcing_dim, device=dev)
    t_c = timeit(lambda: ar_concat(prevs, static, forcing, torch.bfloat16))
    f_in = 2 * F + 4 + info.forcing_dim
    by_c = B * N * (4 * (2 * F + info.forcing_dim) + 2 * 72) + 4 * 4 * N
    rows.append((f"AR concat (fp32 states -> bf16 x, F_in={f_in}, ld 72)", by_c, t_c))
    # external forcings: 5 generated channels of every (sample, step) written into the forcing tensor
    import datetime as dt
    from py4cast_b200 import forcingutils as fu
    H, W = GH, GW
    lat = torch.linspace(55.4, 37.5, H, device=dev).view(H, 1).expand(H, W).contiguous()
    lon = torch.linspace(-12.0, 16.0, W, device=dev).view(1, W).expand(H, W).contiguous()
    fbuf = torch.zeros(B, T, H, W, info.forcing_dim, device=dev)
    dates = [dt.datetime(2023, 6, 1 + b, 6 * (b % 4)) for b in range(B)]
    terms = [dt.timedelta(hours=i + 1) for i in range(T)]
    t_ff = timeit(lambda: fu.fill_external_forcings(fbuf, info.forcing_dim - 5, lat, lon, dates, terms))
    rows.append(("external forcings (5 channels, incl. host date arithmetic)", B * T * N * 5 * 4 + 2 * 4 * N, t_ff))
    scal = fu.external_forcing_scalars(dates, terms).to(dev)
    t_fs = timeit(lambda: fu.fill_external_forcings_from_scalars(fbuf, info.forcing_dim - 5, lat, lon, scal))
    rows.append(("external forcings (5 channels, scalars prepared by the loader)", B * T * N * 5 * 4 + 2 * 4 * N, t_fs))
    # the input path's one-pass assembly: shipped bf16 NWP channels -> fp32 + the five generated channels, whole rows;
    # beside it the two passes it replaces (strided channel copy + fill)
    n_ship = info.forcing_dim - 5
    shipped = torch.randn(B, T, H, W, n_ship, device=dev).to(torch.bfloat16)
    by_a = B * T * N * (2 * n_ship + 4 * info.forcing_dim) + 2 * 4 * N
    t_as = timeit(lambda: fu.assemble_forcings(fbuf, shipped, lat, lon, scal))
    rows.append((f"forcing tensor in one pass ({n_ship} shipped bf16 channels + 5 generated -> fp32 rows)", by_a, t_as))

    def two_pass():
        fbuf[..., :n_ship].copy_(shipped)
        fu.fill_external_forcings_from_scalars(fbuf, n_ship, lat, lon, scal)

    rows.append(("  the two passes it replaces (strided channel copy + fill), same bytes", by_a, timeit(two_pass)))
    # PSD of one prediction step, dataset statistics, un-normalisation
    from py4cast_b200 import dataset_stats as dstat
    from py4cast_b200.metrics import power_spectral_density
    field = torch.randn(B, H, W, F, device=dev)
    t_p = timeit(lambda: power_spectral_density(field))
    rows.append(("PSD profile of (B, H, W, F) fields (column sums + first-row DCT)", 4 * B * N * F, t_p))
    t_s = timeit(lambda: dstat.feature_stats(tgt.detach()))
    rows.append(("feature statistics (nanmean, nanmean of squares, min, max)", 4 * B * T * N * F, t_s))
    nt = NamedTensor(tgt.detach().clone(), names, fn)
    t_u = timeit(lambda: dstat.unnormalize_(nt, info.stats))
    rows.append(("un-normalise predictions in place", 2 * 4 * B * T * N * F, t_u))
    print(f"# element-wise kernels at batch {B}, T={T}, N={N}, F={F}; peak {peak:.0f} GB/s (MEASURED_PEAKS.json or fallback)")
    for label, by, t in rows:
        print(f"{label:60s} {t*1e3:8.3f} ms  {by/1e6:9.1f} MB  {by/t/1e9:7.0f} GB/s  {by/t/1e9/peak:5.2f} of peak")


if __name__ == "__main__":
    main()
