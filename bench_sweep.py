#!/usr/bin/env python
"""Synthetic grid-graph sweep (BASELINE.json configs[4]): forward and forward+backward of a 2-layer
GNNLayer stack (relu(bmm(adj, W x + b)), frozenlake/FrozenLakeNet.py:8-33) on gh x gw grid graphs,
nodes 9-256, hidden 64-256, batch 1k-256k.

Per configuration and precision: time of the 2-layer forward and forward+backward; for the fused tensor-core
layer (bf16x3 / bf16, csrc/azg_grid_tc.cu) the achieved HBM rate on the ALGORITHMIC bytes of a layer
(4H read + 4H written per node) against the measured copy peak, and the issued tensor TFLOP/s; for the
fp32 path (SGEMM + aggregation kernel) the aggregation kernel's rate.  Not the headline metric (bench.py);
a parity-tested (tests/test_gridgnn_gpu.py) roofline table for the graph operator.
usage: python bench_sweep.py [--quick] [--precisions bf16x3,bf16,fp32] [--grids 7x7,8x8] [--hiddens 128] [--batches 65536]
       python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 bench_sweep.py ...
         (graphs are independent: every rank runs `batch` graphs of its own, no collective on the data path; times are
          the max over ranks, graphs/s the whole-job aggregate -- weak scaling)"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from azgnn_b200 import _lib
    from azgnn_b200.gridgnn import GridGNNStack, _GridAggRelu
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--precisions", default="bf16x3,bf16,fp32")
    ap.add_argument("--grids", default="3x3,4x4,6x7,7x7,8x8,12x12,16x16")
    ap.add_argument("--hiddens", default="64,128,256")
    ap.add_argument("--batches", default="1024,16384,262144")
    args = ap.parse_args()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    grids = [tuple(int(v) for v in gs.split("x")) for gs in args.grids.split(",")]
    hiddens = [int(v) for v in args.hiddens.split(",")]
    budget = 2 ** 31  # elements per activation tensor
    rows = []

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
        t = e0.elapsed_time(e1) / reps
        if world > 1:  # max over ranks
            tt = torch.tensor([t], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = tt.item()
        return t

    for gh, gw in grids:
        for H in hiddens:
            n = gh * gw
            for B in ([4096] if args.quick else [int(v) for v in args.batches.split(",")]):
                if B * n * H > budget // 4:
                    continue
                for prec in args.precisions.split(","):
                    torch.manual_seed(rank)
                    net = GridGNNStack(gh, gw, H, layers=2, precision=prec).cuda()
                    x = torch.randn(B, n, H, device="cuda")
                    with torch.no_grad():
                        t_fwd = timed(lambda: net(x))
                        t_agg = None
                        if not net.fused:
                            sup = torch.randn(B, n, H, device="cuda")
                            t_agg = timed(lambda: _GridAggRelu.apply(sup, gh, gw))
                            del sup
                    xg = x.clone().requires_grad_(True)

                    def fb():
                        for p in net.parameters():
                            p.grad = None
                        xg.grad = None
                        net(xg).sum().backward()
                    t_fb = timed(fb, reps=3)
                    layer_bytes = 2 * B * n * H * 4  # a layer reads its input once and writes its output once
                    lin_flops = 2 * B * n * H * H
                    terms = 3 if prec == "bf16x3" else 1
                    r = dict(nodes=n, grid=f"{gh}x{gw}", hidden=H, batch=B, precision=prec, fused=bool(net.fused), fwd_ms=t_fwd,
                             fwd_bwd_ms=t_fb, graphs_per_s_fwd=world * B / t_fwd * 1e3, n_gpus=world)
                    if net.fused:
                        r.update(layer_gbs=2 * layer_bytes / t_fwd / 1e6, frac_hbm=2 * layer_bytes / t_fwd / 1e6 / hbm,
                                 issued_tflops=2 * lin_flops * terms / t_fwd / 1e9)  # per GPU
                    else:
                        r.update(agg_ms=t_agg, layer_gbs=layer_bytes / t_agg / 1e6, frac_hbm=layer_bytes / t_agg / 1e6 / hbm,
                                 issued_tflops=lin_flops * 2 / max(t_fwd - 2 * t_agg, 1e-6) / 1e9)
                    rows.append(r)
                    del net, x, xg
                    torch.cuda.empty_cache()
    if rank != 0:
        dist.destroy_process_group()
        return
    print(f"# grid-graph sweep, 2 layers, {world} GPU(s), batch = graphs per GPU; HBM peak {hbm:.0f} GB/s ({'measured' if peaks else 'fallback'}); fused rows: GB/s = "
          "algorithmic layer bytes / forward time; fp32 rows: GB/s of the aggregation kernel alone, TF/s of the SGEMM alone")
    print(f"{'grid':>6} {'H':>4} {'batch':>7} {'prec':>7} {'fwd ms':>9} {'fwd+bwd ms':>11} {'graphs/s fwd':>13} {'GB/s':>9} {'/HBM':>6} {'TF/s':>8}")
    for r in rows:
        print(f"{r['grid']:>6} {r['hidden']:>4} {r['batch']:>7} {r['precision']:>7} {r['fwd_ms']:>9.3f} {r['fwd_bwd_ms']:>11.3f} "
              f"{r['graphs_per_s_fwd']:>13.3e} {r['layer_gbs']:>9.0f} {r['frac_hbm']:>6.2f} {r['issued_tflops']:>8.1f}")
    print(json.dumps({"metric": "grid_gnn_sweep", "n_gpus": world, "rows": rows}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
