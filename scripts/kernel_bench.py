#!/usr/bin/env python
"""Per-kernel device timing (CUDA events, warm-up, L2 flush between timed launches or
back-to-back) of the sm_100a GEMM family on the shapes of the default MLP and of the
hidden=8192 stress config.  Prints one JSON line per case with achieved bytes/s and the
fraction of the measured HBM copy bandwidth (MEASURED_PEAKS.json)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shallowspeed_b200.ops import cuda as K  # noqa: E402


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "_fallback": True}


def time_fn(fn, iters, flush=None):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / iters
    ts = []
    for _ in range(iters):
        flush.add_(1.0)                      # > L2 sized write
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="default", choices=["default", "stress", "all"])
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"])
    ap.add_argument("--k-splits", type=int, default=0,
                    help="also time the experimental split-K variant of fwd / dgrad (-1 = planner's choice, k >= 2 = force)")
    args = ap.parse_args()
    pk = peaks()
    dev = "cuda"
    flush = torch.zeros(160 * 1024 * 1024 // 4, device=dev)   # 160 MB > 126 MB L2
    shapes = []
    if args.shapes in ("default", "all"):
        shapes += [(32, 784, 128), (32, 128, 127), (32, 123, 10), (4, 784, 128), (128, 784, 128)]
    if args.shapes in ("stress", "all"):
        shapes += [(8, 8192, 8192), (32, 8192, 8192), (128, 8192, 8192)]
    for rows, k, n in shapes:
        x = torch.randn(rows, k, device=dev)
        ld = (k + 1 + 7) // 8 * 8
        Wb = torch.randn(n, ld, device=dev)
        Gb = torch.zeros(n, ld, device=dev)
        W, b = Wb[:, :k], Wb[:, k]
        y = K.empty_padded(rows, n, dev)
        dz = K.empty_padded(rows, n, dev)
        dz.normal_()
        dx = K.empty_padded(rows, k, dev)
        wbytes = n * k * 4
        pr = args.precision
        cases = {
            "fwd": (lambda: K.linear_fwd(x, W, b, relu=True, out=y, precision=pr), wbytes + rows * (k + n) * 4),
            "dgrad": (lambda: K.linear_dgrad(dz, W, mask=x, out=dx, precision=pr), wbytes + rows * (2 * k + n) * 4),
            "wgrad_write": (lambda: K.linear_wgrad(dz, x, Gb[:, :k], accumulate=False, grad_b=Gb[:, k], precision=pr), wbytes + rows * (k + n) * 4),
            "wgrad_acc": (lambda: K.linear_wgrad(dz, x, Gb[:, :k], accumulate=True, grad_b=Gb[:, k], precision=pr), 2 * wbytes + rows * (k + n) * 4),
        }
        if args.k_splits:
            ks = args.k_splits
            cases["fwd_splitk"] = (lambda: K.linear_fwd(x, W, b, relu=True, out=y, precision=pr, k_splits=ks), cases["fwd"][1])
            cases["dgrad_splitk"] = (lambda: K.linear_dgrad(dz, W, mask=x, out=dx, precision=pr, k_splits=ks), cases["dgrad"][1])
        for name, (fn, nbytes) in cases.items():
            warm = time_fn(fn, args.iters)
            cold = time_fn(fn, max(10, args.iters // 3), flush=flush)
            flops = 2.0 * rows * k * n
            print(json.dumps({"kernel": name, "rows": rows, "in": k, "out": n, "us_back_to_back": round(warm, 2),
                              "us_cold_l2": round(cold, 2), "GBps_cold": round(nbytes / cold / 1e3, 1),
                              "frac_hbm_measured": round(nbytes / cold / 1e3 / pk["hbm_gbs"], 3),
                              "tflops_cold": round(flops / cold / 1e6, 2)}), flush=True)


if __name__ == "__main__":
    main()
