"""azgnn_b200 -- B200-native (sm_100a) hot path of the andrpac/alphazero-gnn trainer:
batched MCTS leaf evaluation through the per-game policy/value(+GNN) networks and the
tree operations that feed it, behind the reference's Game / NeuralNet / MCTS / Coach API.

Compute lives in ``csrc/`` (hand-written CUDA, C ABI in ``include/azgnn_b200.h``) and is
loaded by ``_lib``; there is no CPU fallback -- importing the package works anywhere,
calling a compute entry point without the built library or without a GPU raises.
"""
__version__ = "0.1.0"

from . import modules  # noqa: F401
