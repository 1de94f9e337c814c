"""Parameter owners for the B200 path.

These ``nn.Module`` classes hold the weights under the reference's exact parameter
names, shapes, construction order and default initialisers, so that

* ``state_dict()`` / checkpoints are interchangeable with the reference
  (connect4/Connect4GNN.py:199-221 ``{'state_dict', 'gnn'}``), and
* ``torch.manual_seed(s)`` followed by construction yields bit-identical weights to the
  reference classes (which is what lets the golden vectors travel without the weights).

They deliberately have NO torch forward: compute goes through the CUDA library
(``_lib``); calling ``forward`` here raises.  The optimiser step stays in torch.

  Connect4Trunk      <- connect4/Connect4Net.py:11-28
  TicTacToeTrunk     <- tictactoe/TicTacToeNet.py:9-26
  PathGNNLayer       <- gnn_utils.py:5-28          (class GNNLayer)
  PolicyValueGNN     <- gnn_utils.py:87-105
  FrozenLakeGraphNet <- frozenlake/FrozenLakeNet.py:253-295 (class EnhancedNNet)
"""
import torch
import torch.nn as nn


class _ParamsOnly(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} only owns parameters; compute runs in libazgnn_b200.so")


class Connect4Trunk(_ParamsOnly):
    def __init__(self, n, action_size, dropout=0.3):
        super().__init__()
        self.board_x = self.board_y = n
        self.action_size = action_size
        self.conv1 = nn.Conv2d(1, 32, 3, stride=1, padding=1)
        self.conv2 = nn.Conv2d(32, 64, 3, stride=1, padding=1)
        self.fc_policy = nn.Linear(64 * n * n, action_size)
        self.fc_value = nn.Linear(64 * n * n, 1)
        self.dropout = dropout


class TicTacToeTrunk(_ParamsOnly):
    def __init__(self, n, action_size):
        super().__init__()
        self.board_x = self.board_y = n
        self.action_size = action_size
        f = 128 * (n - 2) * (n - 2)
        self.conv1 = nn.Conv2d(1, 32, 3, stride=1, padding=1)
        self.conv2 = nn.Conv2d(32, 64, 3, stride=1, padding=1)
        self.conv3 = nn.Conv2d(64, 128, 3, stride=1)
        self.fc1 = nn.Linear(f, 512)
        self.fc_policy = nn.Linear(512, action_size)
        self.fc2 = nn.Linear(f, 512)
        self.fc_value = nn.Linear(512, 1)


class PathGNNLayer(_ParamsOnly):
    """Attention / update / gate MLPs over [target, source] pairs (gnn_utils.py:11-28)."""

    def __init__(self, feature_dim):
        super().__init__()
        self.feature_dim = feature_dim
        self.attention = nn.Sequential(nn.Linear(feature_dim * 2, 128), nn.ReLU(), nn.Linear(128, 1))
        self.update_net = nn.Sequential(nn.Linear(feature_dim * 2, feature_dim), nn.ReLU(),
                                        nn.Linear(feature_dim, feature_dim))
        self.gate = nn.Sequential(nn.Linear(feature_dim * 2, feature_dim), nn.Sigmoid())


class PolicyValueGNN(_ParamsOnly):
    def __init__(self, feature_dim, num_layers=2):
        super().__init__()
        self.feature_dim = feature_dim
        self.layers = nn.ModuleList([PathGNNLayer(feature_dim) for _ in range(num_layers)])
        self.output_transform = nn.Sequential(nn.Linear(feature_dim, feature_dim), nn.ReLU(),
                                              nn.Linear(feature_dim, feature_dim))


class _GraphConv(_ParamsOnly):
    def __init__(self, i, o):
        super().__init__()
        self.W = nn.Linear(i, o)


class FrozenLakeGraphNet(_ParamsOnly):
    def __init__(self, board_size, action_size, embedding_dim=64, gnn_layers=2):
        super().__init__()
        self.board_x, self.board_y = board_size
        self.action_size = action_size
        self.input_size = self.board_x * self.board_y
        self.embedding_dim = embedding_dim
        self.feature_extractor = nn.Sequential(nn.Linear(self.input_size, 128), nn.ReLU(),
                                               nn.Linear(128, embedding_dim), nn.ReLU())
        self.gnn_layers = nn.ModuleList([_GraphConv(embedding_dim, embedding_dim) for _ in range(gnn_layers)])
        self.policy_head = nn.Linear(embedding_dim, action_size)
        self.value_head = nn.Linear(embedding_dim, 1)
        for m in self.modules():  # FrozenLakeNet.py:289-295
            if isinstance(m, nn.Linear):
                nn.init.xavier_normal_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
