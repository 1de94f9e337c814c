"""Training step of the two-player wrappers on the CUDA library (K3) + the data-parallel
gradient exchange (K5).

    train_two_player   <- Connect4GNNWrapper.train (connect4/Connect4GNN.py:122-197),
                          TicTacToeGNNWrapper.train (tictactoe/TicTacToeGNN.py:89-160),
                          Connect4NNetWrapper.train (connect4/Connect4Net.py:76-108)

Every floating-point operation of forward and backward runs in libazgnn_b200.so; the
`torch.autograd.Function` classes below only route tensors between those kernels, and the
optimizer is torch's Adam exactly as in the reference (re-created per `train` call, lr from
args).  Reference behaviours kept: minibatches are drawn with replacement through the global
NumPy RNG (`np.random.randint`), one std step + one GNN step per "epoch", dropout is live on
the flattened Connect4 features in both steps, and the GNN step only ever updates GNN
parameters (its trunk/head weight gradients are discarded by the next zero_grad in the
reference, so they are not computed here -- SURVEY.md section 8e).

Multi-GPU (one process per GPU, torch.distributed initialised by the caller): rows of the
minibatch are sharded in contiguous chunks; losses are normalised by the GLOBAL batch size;
parameter gradients are summed with one NCCL all-reduce per optimizer.  GNNLayer couples every
row to row 0 (gnn_utils.py:38-74), so the trunk features are all-gathered and the (cheap,
weight-bandwidth bound) layers run replicated, while output_transform + heads + loss shard by
rows; only the rank that owns row 0 contributes layer gradients.
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import ptr, stream
from .mcts import arg


def _lib_check(rc):
    _lib.check(rc)


class _ConvRelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, pad):
        x, w, b = x.contiguous(), w.contiguous(), b.contiguous()
        B, Cin, H, W = x.shape
        Cout = w.shape[0]
        out = torch.empty(B, Cout, H + 2 * pad - 2, W + 2 * pad - 2, dtype=torch.float32, device=x.device)
        _lib_check(_lib.lib().azg_conv3x3_relu_forward(ptr(x), ptr(w), ptr(b), ptr(out), B, Cin, Cout, H, W, pad, stream()))
        ctx.save_for_backward(x, w, out)
        ctx.pad = pad
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w, out = ctx.saved_tensors
        dout = dout.contiguous()
        B, Cin, H, W = x.shape
        Cout = w.shape[0]
        din = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw, db = torch.empty_like(w), torch.empty(Cout, dtype=torch.float32, device=x.device)
        _lib_check(_lib.lib().azg_conv3x3_relu_backward(ptr(x), ptr(w), ptr(out), ptr(dout), ptr(din), ptr(dw), ptr(db), B,
                                                        Cin, Cout, H, W, ctx.pad, stream()))
        return din, dw, db, None


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, relu):
        x, w, b = x.contiguous(), w.contiguous(), b.contiguous()
        M, K = x.shape
        N = w.shape[0]
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        _lib_check(_lib.lib().azg_linear_f32(ptr(x), ptr(w), ptr(b), ptr(y), M, N, K, int(relu), stream()))
        ctx.save_for_backward(x, w, y)
        ctx.relu = int(relu)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = dy.contiguous()
        M, K = x.shape
        N = w.shape[0]
        need_x, need_w, need_b = ctx.needs_input_grad[:3]
        dx = torch.empty_like(x) if need_x else None
        dw = torch.empty_like(w) if need_w else None
        db = torch.empty(N, dtype=torch.float32, device=x.device) if need_b else None
        scratch = torch.empty(M * N, dtype=torch.float32, device=x.device) if ctx.relu else None
        _lib_check(_lib.lib().azg_linear_backward(ptr(dy), ptr(x), ptr(w), ptr(y), M, N, K, ctx.relu, ptr(dx), ptr(dw),
                                                  ptr(db), ptr(scratch), stream()))
        return dx, dw, db, None


class _Mul(torch.autograd.Function):
    """x * mask (dropout with a pre-scaled keep mask); gradient flows to x only."""

    @staticmethod
    def forward(ctx, x, mask):
        x, mask = x.contiguous(), mask.contiguous()
        out = torch.empty_like(x)
        _lib_check(_lib.lib().azg_mul_f32(ptr(x), ptr(mask), ptr(out), x.numel(), stream()))
        ctx.save_for_backward(mask)
        return out

    @staticmethod
    def backward(ctx, dy):
        (mask,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        _lib_check(_lib.lib().azg_mul_f32(ptr(dy), ptr(mask), ptr(dx), dy.numel(), stream()))
        return dx, None


class _PVLoss(torch.autograd.Function):
    """-sum(pi*log_softmax(logits))/norm + sum((v_t - tanh(vraw))^2)/norm, Connect4GNN.py:150-152."""

    @staticmethod
    def forward(ctx, logits, vraw, target_pi, target_v, norm):
        logits, vraw = logits.contiguous(), vraw.contiguous().view(-1)
        B, A = logits.shape
        dev = logits.device
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        logp = torch.empty(B, A, dtype=torch.float32, device=dev)
        v = torch.empty(B, dtype=torch.float32, device=dev)
        dlogits = torch.empty(B, A, dtype=torch.float32, device=dev)
        dvraw = torch.empty(B, dtype=torch.float32, device=dev)
        _lib_check(_lib.lib().azg_policy_value_loss(ptr(logits), ptr(vraw), ptr(target_pi.contiguous()),
                                                    ptr(target_v.contiguous()), B, A, float(norm), ptr(loss), ptr(logp),
                                                    ptr(v), ptr(dlogits), ptr(dvraw), stream()))
        ctx.save_for_backward(dlogits, dvraw)
        ctx.vshape = None
        ctx.mark_non_differentiable(logp, v)
        return loss, logp, v

    @staticmethod
    def backward(ctx, gloss, _glogp, _gv):
        dlogits, dvraw = ctx.saved_tensors
        g = gloss.reshape(())
        return dlogits * g, (dvraw * g).view(-1, 1), None, None, None


class _GraphMeanRelu(torch.autograd.Function):
    """relu(bmm(adj, support)) for FrozenLake's all-ones normalised adjacency (FrozenLakeNet.py:27-33)."""

    @staticmethod
    def forward(ctx, sup, counts):
        sup = sup.contiguous()
        B, _, E = sup.shape
        out = torch.empty_like(sup)
        _lib_check(_lib.lib().azg_graph_mean_relu_forward(ptr(sup), ptr(counts), B, E, ptr(out), stream()))
        ctx.save_for_backward(out, counts)
        return out

    @staticmethod
    def backward(ctx, dout):
        out, counts = ctx.saved_tensors
        dout = dout.contiguous()
        B, _, E = out.shape
        dsup = torch.empty_like(out)
        _lib_check(_lib.lib().azg_graph_mean_relu_backward(ptr(dout), ptr(out), ptr(counts), B, E, ptr(dsup), stream()))
        return dsup, None


def _layer_params(layer):
    return [layer.attention[0].weight, layer.attention[0].bias, layer.attention[2].weight, layer.attention[2].bias,
            layer.update_net[0].weight, layer.update_net[0].bias, layer.update_net[2].weight, layer.update_net[2].bias,
            layer.gate[0].weight, layer.gate[0].bias]


class _GNNLayer(torch.autograd.Function):
    """GNNLayer.forward at B > 1 (gnn_utils.py:38-74): returns the updated target row."""

    @staticmethod
    def forward(ctx, f0, path, *params):
        f0, path = f0.contiguous(), path.contiguous()
        P, F = path.shape
        lib = _lib.lib()
        saved = torch.empty(lib.azg_gnn_layer_saved_floats(P, F), dtype=torch.float32, device=f0.device)
        out0 = torch.empty(F, dtype=torch.float32, device=f0.device)
        cp = _lib.GNNLayerParams(*[ptr(p.contiguous()) for p in params])
        _lib_check(lib.azg_gnn_layer_forward(C.byref(cp), ptr(f0), ptr(path), P, F, ptr(out0), ptr(saved), stream()))
        ctx.save_for_backward(f0, path, saved, *params)
        return out0

    @staticmethod
    def backward(ctx, d_out0):
        f0, path, saved, *params = ctx.saved_tensors
        P, F = path.shape
        lib = _lib.lib()
        d_out0 = d_out0.contiguous()
        grads = [torch.empty_like(p) for p in params]
        d_f0 = torch.empty_like(f0)
        scratch = torch.empty(lib.azg_gnn_layer_scratch_floats(P, F), dtype=torch.float32, device=f0.device)
        cp = _lib.GNNLayerParams(*[ptr(p) for p in params])
        cg = _lib.GNNLayerGrads(*[ptr(g) for g in grads])
        _lib_check(lib.azg_gnn_layer_backward(C.byref(cp), ptr(f0), ptr(path), P, F, ptr(saved), ptr(d_out0), ptr(d_f0),
                                              C.byref(cg), ptr(scratch), stream()))
        return (d_f0, None, *grads)


class CudaOps:
    """The compute vocabulary of the training step, bound to libazgnn_b200.so."""

    @staticmethod
    def conv_relu(x, conv):
        return _ConvRelu.apply(x, conv.weight, conv.bias, int(conv.padding[0]))

    @staticmethod
    def linear(x, lin, relu=False):
        return _Linear.apply(x, lin.weight, lin.bias, relu)

    @staticmethod
    def dropout(x, p):
        if p <= 0:
            return x
        keep = (torch.rand_like(x) >= p).to(torch.float32) / (1.0 - p)  # F.dropout scaling
        return _Mul.apply(x, keep)

    @staticmethod
    def pv_loss(logits, vraw, target_pi, target_v, norm):
        return _PVLoss.apply(logits, vraw, target_pi, target_v, norm)

    @staticmethod
    def gnn_layer(f0, path, layer):
        return _GNNLayer.apply(f0, path, *_layer_params(layer))

    @staticmethod
    def graph_mean_relu(sup, counts):
        return _GraphMeanRelu.apply(sup, counts)

    @staticmethod
    def fl_graph(states, n):
        """K1 graph build (FrozenLakeNet.py:197-213): one-hot node features [B,5,n*n] + node counts [B]"""
        B = states.shape[0]
        nodes = torch.empty(B, 5, n * n, dtype=torch.float32, device=states.device)
        counts = torch.empty(B, dtype=torch.int32, device=states.device)
        _lib_check(_lib.lib().azg_fl_encode_graph(ptr(states), n, B, ptr(nodes), ptr(counts), stream()))
        return nodes, counts


# ----------------------------------------------------------------------------------------- network graphs
def trunk_features(ops, w, boards, training):
    """extract_features: Connect4GNN.py:31-46 (dropout live in training) / TicTacToeGNN.py:25-34."""
    n = w.nnet
    s = boards.view(-1, 1, w.board_x, w.board_y)
    s = ops.conv_relu(s, n.conv1)
    s = ops.conv_relu(s, n.conv2)
    if w.kind == "tictactoe":
        s = ops.conv_relu(s, n.conv3)
    s = s.reshape(s.shape[0], -1)
    if w.kind == "connect4" and training:
        s = ops.dropout(s, float(n.dropout))
    return s


def head_logits(ops, w, feats):
    """apply_policy_value_heads up to the pre-activation outputs (Connect4GNN.py:48-57, TicTacToeGNN.py:36-45)."""
    n = w.nnet
    if w.kind == "connect4":
        return ops.linear(feats, n.fc_policy), ops.linear(feats, n.fc_value)
    return ops.linear(ops.linear(feats, n.fc1, relu=True), n.fc_policy), ops.linear(ops.linear(feats, n.fc2, relu=True), n.fc_value)


def gnn_enhance(ops, gnn, feats):
    """PolicyValueGNN.forward (gnn_utils.py:107-117) on a [B,F] batch: row 0 is the target."""
    x = feats
    if feats.shape[0] > 1:
        f0, path = feats[0], feats[1:]
        for layer in gnn.layers:
            f0 = ops.gnn_layer(f0, path, layer)
        x = torch.cat([f0.unsqueeze(0), path], dim=0)
    ot = gnn.output_transform
    return ops.linear(ops.linear(x, ot[0], relu=True), ot[2])


# ----------------------------------------------------------------------------------------- data parallel
def _world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def shard_rows(B, rank, world):
    """contiguous row range of rank `rank`; row 0 always lives on rank 0"""
    per = (B + world - 1) // world
    lo = min(rank * per, B)
    return lo, min(lo + per, B)


def allreduce_grads(params):
    """Summed all-reduce of the parameter gradients (NCCL over NVLink on GPUs): tensors of >= 1 M elements are
    reduced in place (no staging copies of the 39 MB weight gradients), the small ones travel in one flat bucket."""
    rank, world = _world()
    if world == 1:
        return
    params = [p for p in params]
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    big = [p for p in params if p.grad.numel() >= (1 << 20) and p.grad.is_contiguous()]
    small = [p for p in params if not (p.grad.numel() >= (1 << 20) and p.grad.is_contiguous())]
    # inside a CUDA-graph capture (the data-parallel step is captured with its collectives) the reductions are enqueued
    # synchronously with respect to the capturing stream; eager steps overlap them with the small bucket's packing
    capturing = p.grad.is_cuda and torch.cuda.is_current_stream_capturing()
    works = [dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, async_op=not capturing) for p in big]
    if capturing:
        works = []
    if small:
        flat = torch._utils._flatten_dense_tensors([p.grad for p in small])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        for p, g in zip(small, torch._utils._unflatten_dense_tensors(flat, [p.grad for p in small])):
            p.grad.copy_(g)
    for wk in works:
        wk.wait()


class _ShareRow0Grad(torch.autograd.Function):
    """Identity on the updated target row; its backward sums d(loss)/d(row 0) over ranks.  Only the rank that owns
    row 0 has a non-zero contribution, so afterwards EVERY rank holds the true gradient and runs the GNNLayer
    backward itself: the 100 M layer gradients are recomputed (0.35 ms of weight streaming) instead of being
    moved (400 MB all-reduce), and stay bit-identical across ranks."""

    @staticmethod
    def forward(ctx, f0):
        return f0.view_as(f0)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        return g


def _overlap_hooks(w):
    """Multi-rank GNN step: the output_transform gradients are complete BEFORE the backward pass reaches the GNNLayers
    (0.35 ms of weight streaming, replicated on every rank).  Post-accumulate hooks start their all-reduce at that
    moment on NCCL's stream, so the 79 MB exchange runs under the layer backward instead of after it; `exchange_grads`
    then only waits.  In a captured step the two become parallel branches of the graph."""
    if getattr(w, "_ot_hooks", None) is not None:
        return
    w._ot_pending = []
    w._ot_hooks_on = False

    def hook(p):
        if w._ot_hooks_on and p.grad is not None and p.grad.numel() >= (1 << 20) and p.grad.is_contiguous():
            w._ot_pending.append((p, dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, async_op=True)))
    w._ot_hooks = [p.register_post_accumulate_grad_hook(hook) for p in w.gnn.output_transform.parameters()]


def exchange_grads(w, which):
    """The one exchange of a data-parallel step.  "std": every trunk/head gradient.
    "gnn": the row-sharded output_transform gradients only -- the GNNLayer gradients are already complete and
    identical on every rank (see _ShareRow0Grad); the large ones were started by `_overlap_hooks` during backward."""
    if which == "std":
        allreduce_grads(list(w.nnet.parameters()))
        return
    pending = getattr(w, "_ot_pending", None) or []
    done = set()
    for p, work in pending:
        work.wait()
        done.add(id(p))
    if pending:
        w._ot_pending = []
    allreduce_grads([p for p in w.gnn.output_transform.parameters() if id(p) not in done])


def std_step(ops, w, boards, target_pi, target_v):
    """Loss of the standard-network step on this rank's rows (global normalisation)."""
    rank, world = _world()
    B = boards.shape[0]
    lo, hi = shard_rows(B, rank, world)
    if hi <= lo:
        return None
    feats = trunk_features(ops, w, boards[lo:hi], training=True)
    logits, vraw = head_logits(ops, w, feats)
    loss, _, _ = ops.pv_loss(logits, vraw, target_pi[lo:hi], target_v[lo:hi], B)
    return loss


def gnn_step(ops, w, boards, target_pi, target_v):
    """Loss of the GNN step on this rank's rows; gradients reach GNN parameters only."""
    rank, world = _world()
    B = boards.shape[0]
    lo, hi = shard_rows(B, rank, world)
    with torch.no_grad():  # trunk/head weight gradients of this step are never consumed
        mine = trunk_features(ops, w, boards[lo:hi], training=True) if hi > lo else boards.new_zeros(0, w.feature_dim)
        if world > 1:
            per = (B + world - 1) // world
            pad = torch.zeros(per, w.feature_dim, dtype=mine.dtype, device=mine.device)
            pad[:mine.shape[0]] = mine
            parts = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(parts, pad)
            feats = torch.cat(parts, dim=0)[:B]
        else:
            feats = mine
    gnn = w.gnn
    x = feats
    f0 = None
    if B > 1:
        f0, path = feats[0], feats[1:]
        for layer in gnn.layers:
            f0 = ops.gnn_layer(f0, path, layer)
        if world > 1:
            f0 = _ShareRow0Grad.apply(f0)
        x = torch.cat([f0.unsqueeze(0), path], dim=0)
    # ranks that do not own row 0 still take part in its gradient exchange (with a zero contribution)
    anchor = (f0 * 0.0).sum().reshape(1) if (world > 1 and f0 is not None and rank != 0) else None
    if hi <= lo:
        return anchor
    ot = gnn.output_transform
    enh = ops.linear(ops.linear(x[lo:hi], ot[0], relu=True), ot[2])
    heads = _FrozenHeads(w)
    logits, vraw = head_logits(ops, heads, enh)
    loss, _, _ = ops.pv_loss(logits, vraw, target_pi[lo:hi], target_v[lo:hi], B)
    return loss if anchor is None else loss + anchor


class _FrozenHeads:
    """view of a wrapper whose head parameters are detached: the GNN step needs d(loss)/d(enhanced)
    through the heads but no head-weight gradients"""

    def __init__(self, w):
        self.kind = w.kind

        class _L:
            def __init__(self, lin):
                self.weight, self.bias = lin.weight.detach(), lin.bias.detach()
        n = w.nnet
        names = ["fc_policy", "fc_value"] + (["fc1", "fc2"] if w.kind == "tictactoe" else [])
        self.nnet = type("N", (), {k: _L(getattr(n, k)) for k in names})()


def _shared_rng():
    """Multi-rank train() call: ONE broadcast of a seed drawn from rank 0's global NumPy RNG; every rank then draws the same
    minibatch indices locally (a broadcast per minibatch was a host synchronisation per optimizer step)."""
    seed = torch.tensor([np.random.randint(0, 2 ** 31 - 1)], dtype=torch.int64)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    seed = seed.to(dev)
    dist.broadcast(seed, src=0)
    return np.random.RandomState(int(seed.item()))


def _sample(examples, batch_size, rng=None):
    """np.random.randint minibatch with replacement (Connect4GNN.py:141,160); with several ranks every rank draws from
    the call's shared stream (`_shared_rng`) so that all work on the same minibatch."""
    if rng is not None:
        return rng.randint(0, len(examples), min(len(examples), batch_size))
    idx = np.random.randint(0, len(examples), min(len(examples), batch_size))
    rank, world = _world()
    if world > 1:
        t = torch.as_tensor(idx, dtype=torch.int64)
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = t.to(dev)
        dist.broadcast(t, src=0)
        idx = t.cpu().numpy()
    return idx


def _backward(w, which, loss):
    """loss.backward() with the overlapped output_transform exchange armed for multi-rank NCCL GNN steps"""
    if loss is None:
        return
    _rank, world = _world()
    overlap = which == "gnn" and world > 1 and dist.get_backend() == "nccl" and getattr(w, "gnn", None) is not None
    if overlap:
        _overlap_hooks(w)
        w._ot_hooks_on = True
    try:
        loss.backward()
    finally:
        if overlap:
            w._ot_hooks_on = False


class _GraphedStep:
    """One optimizer step (zero_grad -> loss -> backward -> Adam.step) of a fixed batch shape, captured in a CUDA
    graph: the ~120 launches of a step are replayed by one `cudaGraphLaunch` instead of being issued from Python
    (the step is launch-bound otherwise: 3.1 ms eager vs 1.9 ms of kernel time per epoch)."""

    def __init__(self, step_fn, w, params, opt, shapes, which="std"):
        dev = w.device
        self.static = [torch.zeros(s, dtype=torch.float32, device=dev) for s in shapes]
        self.graph = None
        self.step_fn, self.w, self.params, self.opt, self.which = step_fn, w, params, opt, which

    def _body(self):
        self.opt.zero_grad(set_to_none=True)
        loss = self.step_fn(CudaOps, self.w, *self.static)
        _backward(self.w, self.which, loss)
        exchange_grads(self.w, self.which)  # no-op on one rank; NCCL nodes of the captured graph otherwise
        self.opt.step()

    def run(self, tensors, eager):
        for dst, src in zip(self.static, tensors):
            dst.copy_(src, non_blocking=True)
        if eager:  # first step after the optimizer was created: Adam allocates its state here
            self._body()
            return
        if self.graph is None:
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            self.opt.zero_grad(set_to_none=True)
            # thread_local: NCCL's watchdog thread may query events while this thread captures
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self._body()
        self.graph.replay()


def _reset_adam(opt):
    """fresh-Adam semantics (the reference re-creates its optimizers on every train call) for a cached optimizer"""
    for st in opt.state.values():
        for v in st.values():
            if torch.is_tensor(v):
                v.zero_()


def release_captured_steps(w):
    """Drop the CUDA graphs (and cached optimizers) of a wrapper.  Multi-rank graphs hold NCCL kernels: NCCL requires
    them to be destroyed before the communicator is (`destroy_process_group` otherwise blocks forever)."""
    import gc
    cache = w.__dict__.pop("_train_cache", None)
    if cache:
        for entry in cache.values():
            if isinstance(entry, dict):
                for step in entry.get("steps", {}).values():
                    step.graph = None
                entry.get("steps", {}).clear()
        cache.clear()
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def _train_cache(w, key, builder):
    cache = w.__dict__.setdefault("_train_cache", {})
    if key not in cache:
        cache[key] = builder()
    return cache[key]


def train_two_player(w, examples, gnn_examples=None, ops=CudaOps):
    lr = float(arg(w.args, "lr"))
    epochs, batch_size = int(arg(w.args, "epochs")), int(arg(w.args, "batch_size"))
    dev = w.device
    rank, world = _world()
    nnet_params = list(w.nnet.parameters())
    gnn_params = list(w.gnn.parameters()) if getattr(w, "gnn", None) is not None else []
    # torch's own Adam, as in the reference; its single-kernel (`fused`) implementation moves each of p, g, m, v once
    # instead of once per foreach pass (1.8 ms -> 0.8 ms for the 120 M GNN parameters)
    # multi-rank steps are captured too, NCCL collectives included (feature all-gather, row-0 gradient and weight-gradient
    # all-reduces become graph nodes): eager 8-rank epochs were launch-bound, 3.7 ms against 2.2 ms on one GPU
    graphed = dev.type == "cuda" and ops is CudaOps and not os.environ.get("AZG_TRAIN_EAGER") and \
        (world == 1 or (dist.get_backend() == "nccl" and not os.environ.get("AZG_TRAIN_DP_EAGER")))
    if graphed:
        # optimizers (state zeroed = re-created) and captured steps persist across train() calls
        def build():
            o1 = torch.optim.Adam(nnet_params, lr=lr, fused=True, capturable=True)
            o2 = torch.optim.Adam(gnn_params, lr=lr, fused=True, capturable=True) if gnn_params else None
            return {"nnet_opt": o1, "gnn_opt": o2, "steps": {}, "warm": set()}
        cache = _train_cache(w, ("opt", lr), build)
        nnet_opt, gnn_opt = cache["nnet_opt"], cache["gnn_opt"]
        _reset_adam(nnet_opt)
        if gnn_opt is not None:
            _reset_adam(gnn_opt)
    else:
        kw = {"fused": True} if dev.type == "cuda" else {}
        nnet_opt = torch.optim.Adam(nnet_params, lr=lr, **kw)
        gnn_opt = torch.optim.Adam(gnn_params, lr=lr, **kw) if gnn_params else None

    def run_step(which, step_fn, params, opt, tensors):
        if not graphed:
            opt.zero_grad()
            loss = step_fn(ops, w, *tensors)
            _backward(w, which, loss)
            exchange_grads(w, which)
            opt.step()
            return
        key = (which, tuple(tuple(t.shape) for t in tensors))
        if key not in cache["steps"]:
            if len(cache["steps"]) >= 6:  # every capture pins its own gradient pool (0.5 GB for the GNN step): odd batch
                opt.zero_grad()           # shapes beyond a handful run eagerly instead of being captured
                loss = step_fn(ops, w, *tensors)
                _backward(w, which, loss)
                exchange_grads(w, which)
                opt.step()
                cache["warm"].add(which)
                return
            cache["steps"][key] = _GraphedStep(step_fn, w, params, opt, [t.shape for t in tensors], which)
        eager = which not in cache["warm"]
        cache["steps"][key].run(tensors, eager)
        cache["warm"].add(which)

    shared = _shared_rng() if world > 1 else None
    for _ in range(epochs):
        if examples is not None and len(examples) > 0:
            idx = _sample(examples, batch_size, shared)
            if hasattr(examples, "sample"):  # replay.DeviceExamples: the minibatch is gathered on the device
                boards, target_pis, target_vs = examples.sample(batch_size, idx=idx)
            else:
                boards, pis, vs = list(zip(*[examples[i] for i in idx]))
                boards = torch.FloatTensor(np.array(boards)).to(dev)
                target_pis = torch.FloatTensor(np.array(pis)).to(dev)
                target_vs = torch.FloatTensor(np.array(vs).astype(np.float64)).to(dev)
            run_step("std", std_step, nnet_params, nnet_opt, (boards, target_pis, target_vs))
        if gnn_opt is not None and gnn_examples is not None and len(gnn_examples) > 0:
            idx = _sample(gnn_examples, batch_size, shared)
            if hasattr(gnn_examples, "sample"):  # replay.DeviceGnnExamples
                boards, expanded_pis, expanded_vs = gnn_examples.sample(batch_size, idx=idx)
            else:
                batch = [gnn_examples[i] for i in idx]
                boards = torch.FloatTensor(np.array([b[0] for b in batch])).to(dev)
                expanded_pis = torch.FloatTensor(np.array([b[4] for b in batch])).to(dev)
                expanded_vs = torch.FloatTensor(np.array([b[5] for b in batch]).astype(np.float64)).to(dev)
            run_step("gnn", gnn_step, gnn_params, gnn_opt, (boards, expanded_pis, expanded_vs))
    w.weights_changed()


def fl_step(ops, w, states, target_pi, target_v):
    """Loss of one FrozenLake minibatch (FrozenLakeNet.py:113-160) with all graphs of the batch evaluated
    together: nodes [B,5,n^2] -> feature_extractor -> L x (Linear, aggregate) -> node 0 -> heads -> loss.
    The reference's `log(out_pi.clamp(min=1e-8))` equals log_softmax unless a probability underflows 1e-8."""
    net = w.nnet
    n = w.n
    nodes, counts = ops.fl_graph(states, n)
    B = nodes.shape[0]
    x = nodes.reshape(B * 5, n * n)
    fe = net.feature_extractor
    x = ops.linear(ops.linear(x, fe[0], relu=True), fe[2], relu=True)
    for layer in net.gnn_layers:
        sup = ops.linear(x, layer.W)
        x = ops.graph_mean_relu(sup.reshape(B, 5, -1), counts).reshape(B * 5, -1)
    cur = x.reshape(B, 5, -1)[:, 0, :]
    logits, vraw = ops.linear(cur, net.policy_head), ops.linear(cur, net.value_head)
    loss, _, _ = ops.pv_loss(logits, vraw, target_pi, target_v, B)
    return loss


def train_frozenlake(w, examples, ops=CudaOps):
    """FrozenLakeNet.train (FrozenLakeNet.py:76-176): fresh Adam, `epochs` passes over shuffled minibatches,
    NaN batches skipped, gradient norm clipped to 1.0.  Graph construction and forward/backward run in the
    CUDA library for the whole minibatch at once instead of one Python iteration per board."""
    from .mcts import pack_states
    if not examples or len(examples) < 4:
        print("Not enough examples for training, need at least 4")
        return
    print(f"Training on {len(examples)} examples")
    lr, epochs, batch_size = float(arg(w.args, "lr")), int(arg(w.args, "epochs")), int(arg(w.args, "batch_size"))
    params = list(w.nnet.parameters())
    w.optimizer = torch.optim.Adam(params, lr=lr)
    examples = [(x[0], x[1], x[2]) for x in examples if x[2] is not None]
    for epoch in range(epochs):
        np.random.shuffle(examples)
        bs = min(len(examples), batch_size)
        epoch_loss, num_batches = 0.0, 0
        for i in range(0, len(examples), bs):
            boards, pis, vs = list(zip(*examples[i:i + bs]))
            if any(np.isnan(np.sum(x)) for x in boards) or any(np.isnan(np.sum(x)) for x in pis) or any(np.isnan(x) for x in vs):
                print("NaN values detected in input data, skipping batch")
                continue
            states = torch.as_tensor(pack_states("frozenlake", np.array(boards))).to(w.device)
            target_pis = torch.FloatTensor(np.array(pis)).to(w.device)
            target_vs = torch.FloatTensor(np.array(vs).astype(np.float64)).to(w.device)
            w.optimizer.zero_grad()
            loss = fl_step(ops, w, states, target_pis, target_vs)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, max_norm=1.0)
            w.optimizer.step()
            epoch_loss += loss.item()
            num_batches += 1
        if num_batches > 0:
            print(f"Epoch {epoch + 1}/{epochs} - Loss: {epoch_loss / num_batches:.4f}")
    w.weights_changed()
