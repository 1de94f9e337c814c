"""The tensor-core splits on a TRAINED network (the 1e-5 margin was only ever shown on random-init weights).

Self-play with the random-init network collects examples, `NeuralNet.train` runs 200 epochs (200 standard + 200 GNN
optimizer steps, Connect4GNN.py:135-197) on them, and 4,096 positions (half random cell fills, half positions reached in
self-play) are then evaluated by every precision and compared with the oracle (torch fp32 on the CPU, the reference's
arithmetic).  Contract: pi and v within 1e-5 for whatever `b200_precision: auto` picked; the errors of the individual
modes are printed (the guard, not a fixed mode, is what holds the contract when weights grow)."""
import numpy as np
import pytest
import torch

from oracle import nets as onets

from azgnn_b200 import _lib, games
from azgnn_b200.nets import B200Connect4GNNWrapper
from azgnn_b200.selfplay import BatchedSelfPlay, unpack_boards
from helpers import dotdict

pytestmark = pytest.mark.gpu


def test_splits_hold_the_contract_on_a_trained_checkpoint():
    n = 7
    game = games.Connect4Game(n)
    args = dotdict(dict(lr=1e-3, dropout=0.3, epochs=200, batch_size=64, gnn_layers=2, use_gnn=True, numMCTSSims=10, cpuct=1.0,
                        expand_by=5, tempThreshold=15))
    torch.manual_seed(0)
    w = B200Connect4GNNWrapper(game, args)
    assert w.precision == _lib.PREC_AUTO
    before = {k: v.detach().clone() for k, v in w.gnn.output_transform.state_dict().items()}
    sp = BatchedSelfPlay(game, w, args, 256, seed=3, collect_examples="device")
    sp.play(256)
    std, gnn = sp.device_examples, sp.device_gnn_examples
    assert len(std) > 2000 and len(gnn) > 1000
    np.random.seed(0)
    w.train(std.shuffled(), gnn.shuffled())
    moved = max(float((v - before[k]).abs().max()) for k, v in w.gnn.output_transform.state_dict().items())
    assert moved > 1e-3  # 200 Adam steps at lr 1e-3 moved the contraction weights

    rng = np.random.default_rng(99)
    rand_boards = rng.integers(-1, 2, size=(2048, n, n)).astype(np.int64)
    sp2 = BatchedSelfPlay(game, w, args, 512, seed=4, collect_examples=False)
    seen = []
    for _ in range(4):  # positions the trained network reaches itself (canonical boards of the roots after 3..12 plies)
        for _ in range(3):
            sp2.step_all()
        seen.append(unpack_boards("connect4", n, sp2.mcts.arena.to_host(sp2.mcts.arena.get_roots())))
    boards = np.concatenate([rand_boards] + seen).astype(np.int64)
    assert boards.shape[0] == 4096
    p = {k: v.detach().cpu() for k, v in w.nnet.state_dict().items()}
    q = {k: v.detach().cpu() for k, v in w.gnn.state_dict().items()}
    bt = onets.boards_to_tensor(boards)
    with torch.no_grad():
        spi, sv = onets.c4_predict(p, bt, n)
        gpi, gv = onets.c4_predict_with_gnn(p, q, bt, n)
    want = {"pi": spi, "v": sv, "pi_gnn": gpi, "v_gnn": gv}
    states = w.states_from_boards(boards)
    both = _lib.EVAL_STD | _lib.EVAL_GNN

    def errors(prec, fold=False):
        w.fold_heads = fold
        o = w.forward_states(states, both, precision=prec)
        w.fold_heads = False
        return {k: float((o[k].cpu().reshape(t.shape) - t).abs().max()) for k, t in want.items()}
    print(f"trained checkpoint: max |W0| {float(w.gnn.output_transform[0].weight.detach().abs().max()):.4f}, max policy prob {float(gpi.max()):.3f}, "
          f"max |v| {float(gv.abs().max()):.3f}")
    table = {}
    for name, prec, fold in (("fp32", _lib.PREC_FP32, False), ("bf16x3", _lib.PREC_BF16X3, False), ("f16f8", _lib.PREC_F16F8, False),
                             ("f16f8+fold", _lib.PREC_F16F8, True), ("f16f8ks", _lib.PREC_F16F8_KS, False),
                             ("f16f8ks+fold", _lib.PREC_F16F8_KS, True), ("bf16x3ks", _lib.PREC_BF16X3_KS, False),
                             ("bf16", _lib.PREC_BF16, False)):
        table[name] = errors(prec, fold)
        print(f"  {name:11s} max |d pi| {table[name]['pi']:.2e}  |d v| {table[name]['v']:.2e}  |d pi_gnn| {table[name]['pi_gnn']:.2e}  "
              f"|d v_gnn| {table[name]['v_gnn']:.2e}")
    chosen = _lib.PRECISION_NAMES[w.active_precision()]
    print(f"  auto -> {chosen}; probe report {w.precision_report}")
    auto = errors(None)
    # (1) what the default configuration delivers: the contract, whatever mode the guard had to fall back to
    assert max(auto.values()) <= 1e-5, (chosen, auto)
    assert max(table["fp32"].values()) <= 1e-5
    # (2) the tensor-core splits on a trained network: the tensor core's truncating fp32 accumulation (~1e-5 at K = 3136,
    # DESIGN.md section 4) puts them AT the 1e-5 line (8e-6 .. 1.3e-5 over several runs), so a fixed mode cannot promise the
    # fp32 contract here; their stated tolerance on trained weights is 5e-5, and a mode the guard accepts (probe <= AUTO_TOL)
    # must be inside 1e-5
    for name in ("bf16x3", "f16f8", "f16f8+fold", "f16f8ks", "f16f8ks+fold"):
        assert max(table[name].values()) <= 5e-5, (name, table[name])
    # (3) the K-split accumulation (four accumulators of K/4 each, summed with round-to-nearest adds) removes most of that
    # floor: inside the contract with margin on the trained network
    # (measured over three checkpoints: 6.6e-6 .. 7.3e-6 against 9.1e-6 .. 1.04e-5 in one accumulator; what is left is the
    # 16-17 bit operand representation of the trunk's bf16 split and of the fp16+FP8 split, not the accumulation)
    # (ratio to the single accumulator over eight checkpoints: 0.61 .. 0.83; asserted loosely, every run trains a new network)
    assert max(table["f16f8ks"].values()) <= max(table["f16f8"].values()), (table["f16f8ks"], table["f16f8"])
    assert max(table["f16f8ks"].values()) <= 1e-5, table["f16f8ks"]
    for name in ("bf16x3", "f16f8", "f16f8ks", "bf16x3ks"):
        if w.precision_report.get(name, 1.0) <= w.AUTO_TOL:
            assert max(table[name].values()) <= 1e-5, (name, table[name])
