#!/usr/bin/env python
"""Synthetic grid-graph sweep (BASELINE.json configs[4]): forward and forward+backward of a 2-layer
GNNLayer stack (relu(bmm(adj, W x + b)), frozenlake/FrozenLakeNet.py:8-33) on gh x gw grid graphs,
nodes 9-256, hidden 64-256, batch 1k-256k.  Reports per-layer-pass time, achieved GB/s of the
aggregation kernel against the measured HBM copy peak and TFLOP/s of the dense Linear (fp32 SGEMM).
Not the headline metric (bench.py); a parity-tested roofline table for the graph operator.
usage: python bench_sweep.py [--quick]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    import torch
    from azgnn_b200 import _lib
    from azgnn_b200.gridgnn import GridGNNStack, _GridAggRelu
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    grids = [(3, 3), (4, 4), (6, 7), (7, 7), (8, 8), (16, 16)]
    hiddens = [64, 128, 256]
    budget = 2 ** 31  # elements per activation tensor
    rows = []
    for gh, gw in grids:
        for H in hiddens:
            n = gh * gw
            for B in ([4096] if args.quick else [1024, 16384, 262144]):
                if B * n * H > budget // 4:
                    continue
                torch.manual_seed(0)
                net = GridGNNStack(gh, gw, H, layers=2).cuda()
                x = torch.randn(B, n, H, device="cuda")

                def timed(fn, reps=5):
                    for _ in range(2):
                        fn()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(reps):
                        fn()
                    e1.record()
                    torch.cuda.synchronize()
                    return e0.elapsed_time(e1) / reps

                with torch.no_grad():
                    t_fwd = timed(lambda: net(x))
                    sup = torch.randn(B, n, H, device="cuda")
                    t_agg = timed(lambda: _GridAggRelu.apply(sup, gh, gw))
                xg = x.clone().requires_grad_(True)

                def fb():
                    for p in net.parameters():
                        p.grad = None
                    xg.grad = None
                    net(xg).sum().backward()
                t_fb = timed(fb, reps=3)
                agg_bytes = 2 * B * n * H * 4  # read support once, write output once
                lin_flops = 2 * B * n * H * H
                rows.append(dict(nodes=n, grid=f"{gh}x{gw}", hidden=H, batch=B, fwd_ms=t_fwd, fwd_bwd_ms=t_fb, agg_ms=t_agg,
                                 agg_gbs=agg_bytes / t_agg / 1e6, agg_frac_hbm=agg_bytes / t_agg / 1e6 / hbm,
                                 graphs_per_s_fwd=B / t_fwd * 1e3,
                                 linear_tflops=lin_flops * 2 / max(t_fwd - 2 * t_agg, 1e-6) / 1e9))
                del net, x, sup, xg
                torch.cuda.empty_cache()
    print(f"# grid-graph sweep, 2 layers, fp32; HBM peak {hbm:.0f} GB/s ({'measured' if peaks else 'fallback'})")
    print(f"{'grid':>6} {'H':>4} {'batch':>7} {'fwd ms':>9} {'fwd+bwd ms':>11} {'graphs/s fwd':>13} {'agg GB/s':>9} {'agg/HBM':>8} {'lin TF/s':>9}")
    for r in rows:
        print(f"{r['grid']:>6} {r['hidden']:>4} {r['batch']:>7} {r['fwd_ms']:>9.3f} {r['fwd_bwd_ms']:>11.3f} {r['graphs_per_s_fwd']:>13.3e} "
              f"{r['agg_gbs']:>9.0f} {r['agg_frac_hbm']:>8.2f} {r['linear_tflops']:>9.1f}")
    print(json.dumps({"metric": "grid_gnn_sweep", "rows": rows}))


if __name__ == "__main__":
    main()
