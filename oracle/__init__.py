"""oracle/ -- TEST INFRASTRUCTURE ONLY.

A CPU restatement of the reference algorithm (andrpac/alphazero-gnn) for the
hot path this repository accelerates.  It exists so the CUDA path can be
checked for parity on machines where /root/reference is absent.

Rules of use
------------
* Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
  ``cpu_baseline`` / ``--impl reference`` legs may import anything from here,
  and only as the checker / the timed CPU baseline -- never as the thing
  shipped.  Nothing under ``alphazero-gnn_b200/`` imports ``oracle``.
* Parity pinning: the reference ships no tests or golden vectors
  (SURVEY.md section 4), so the oracle is pinned against outputs of the
  reference itself, generated in the build container by
  ``tests/golden/make_golden.py`` (which imports /root/reference read-only) and
  committed as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks
  every function here against those vectors.

Modules
-------
rules   numpy restatement of the three games' rules       (Connect4Game.py,
        TicTacToeGame.py, FrozenLakeGame.py)
nets    torch-fp32 functional restatement of the networks (Connect4Net.py,
        Connect4GNN.py, TicTacToe*.py, FrozenLakeNet.py, gnn_utils.py)
mcts    restatement of MCTS.py with an explicit stack (no recursion)
"""
