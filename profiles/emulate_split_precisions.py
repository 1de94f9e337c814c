"""CPU emulation of operand-rounding schemes for the two F x F contractions (exact products, fp64 accumulation):
max |d pi|, max |d v| against the fp32 module for bf16x3 (shipped), fp16 + two FP8 correction products,
an fp16 two-term split and plain bf16, with every weight matrix scaled by 0.3 / 1 / 3 / 10.  No GPU needed."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import nets as onets
from azgnn_b200 import modules
torch.manual_seed(0)
n=7
def run(scale):
    torch.manual_seed(0)
    nnet = modules.Connect4Trunk(n, n+1); gnn = modules.PolicyValueGNN(64*n*n, 2)
    with torch.no_grad():
        for p_ in list(nnet.parameters()) + list(gnn.output_transform.parameters()):
            if p_.dim() > 1: p_.mul_(scale ** 0.5)
    p, q = dict(nnet.state_dict()), dict(gnn.state_dict())
    boards = np.random.default_rng(11).integers(-1, 2, size=(512, n, n)).astype(np.int64)
    bt = onets.boards_to_tensor(boards)
    with torch.no_grad():
        gpi, gv = onets.c4_predict_with_gnn(p, q, bt, n)
        # features
        x = bt.view(-1,1,n,n)
        h = torch.relu(torch.nn.functional.conv2d(x, p['conv1.weight'], p['conv1.bias'], padding=1))
        h = torch.relu(torch.nn.functional.conv2d(h, p['conv2.weight'], p['conv2.bias'], padding=1))
        feat = h.reshape(h.shape[0], -1)
    W0, b0, W2, b2 = q['output_transform.0.weight'], q['output_transform.0.bias'], q['output_transform.2.weight'], q['output_transform.2.bias']
    def heads(E):
        logits = E @ p['fc_policy.weight'].T + p['fc_policy.bias']
        v = torch.tanh(E @ p['fc_value.weight'].T + p['fc_value.bias'])
        return torch.softmax(logits, 1), v
    def mm64(a, b):  # exact products, fp64 accumulate (tensor core accumulates in fp32: add that rounding noise separately)
        return (a.double() @ b.double().T)
    def bf16x3(X, W):
        Xh = X.bfloat16().float(); Xl = (X - Xh).bfloat16().float()
        Wh = W.bfloat16().float(); Wl = (W - Wh).bfloat16().float()
        return (mm64(Xh, Wh) + mm64(Xh, Wl) + mm64(Xl, Wh)).float()
    def q8(t):
        return t.to(torch.float8_e4m3fn).float()
    def f16f8(X, W, s=2.0**12):
        Xh = X.half().float(); Wh = W.half().float()
        Xl8 = q8(((X - Xh) * s).clamp(-448, 448)); Wl8 = q8(((W - Wh) * s).clamp(-448, 448))
        X8 = q8(X.clamp(-448,448)); W8 = q8(W.clamp(-448,448))
        return (mm64(Xh, Wh) + (mm64(X8, Wl8) + mm64(Xl8, W8)) / s).float()
    def f16x2(X, W):
        Xh = X.half().float(); Xl=(X-Xh).half().float(); Wh = W.half().float()
        return (mm64(Xh, Wh) + mm64(Xl, Wh)).float()
    def bf16(X, W):
        return mm64(X.bfloat16().float(), W.bfloat16().float()).float()
    res = {}
    for name, f in (('bf16x3', bf16x3), ('f16+f8x2', f16f8), ('f16x2', f16x2), ('bf16', bf16)):
        with torch.no_grad():
            H = torch.relu(f(feat, W0) + b0)
            E = f(H, W2) + b2
            pi, v = heads(E)
        res[name] = (float((pi - gpi).abs().max()), float((v.squeeze() - gv.squeeze()).abs().max()))
    # fp32 exact reference vs fp64 for context
    print('scale', scale, {k: (f"{a:.2e}", f"{b:.2e}") for k, (a, b) in res.items()}, 'max|feat|', float(feat.abs().max()), 'max|W0|', float(W0.abs().max()))
for sc in (0.3, 1.0, 3.0, 10.0):
    run(sc)
