"""oracle.selfplay -- TEST INFRASTRUCTURE: restatement of Coach.executeEpisode (Coach.py:27-79)
for one sequential game, used as the CPU baseline of the self-play moves/s metric."""
import numpy as np


def execute_episode(game, mcts, args, max_moves=10_000):
    """Plays one episode; returns (number of moves, result).  Examples are formed as the reference
    does (symmetries + expand_tree when use_gnn) so the timed work is the same."""
    train_examples, gnn_examples = [], []
    board, cur, step = game.getInitBoard(), 1, 0
    use_gnn = hasattr(args, 'use_gnn') and args.use_gnn
    while True:
        step += 1
        canon = game.getCanonicalForm(board, cur)
        temp = int(step < args.tempThreshold)
        pi = mcts.getActionProb(canon, temp=temp)
        sym = game.getSymmetries(canon, pi) if hasattr(game, "getSymmetries") else [(canon, pi)]
        for b, p in sym:
            train_examples.append([b, cur, p, None])
        if use_gnn:
            for s, rec in mcts.expand_tree(canon, expand_by=getattr(args, 'expand_by', 5)).items():
                gnn_examples.append([canon, cur, *rec, None])
        action = np.random.choice(len(pi), p=pi)
        board, cur = game.getNextState(board, cur, action)
        r = game.getGameEnded(board, cur)
        if r != 0 or step >= max_moves:
            return step, r
