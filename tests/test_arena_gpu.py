"""Arena parity on the GPU: the same cases as tests/test_arena_host.py, through libazgnn_b200.so."""
import numpy as np
import pytest
import torch

import arena_cases as cases
from azgnn_b200 import _lib
from azgnn_b200.arena import DeviceArena, GAME_KINDS, action_size

pytestmark = pytest.mark.gpu


def make_arena(kind, n, n_games, sims, cpuct, **kw):
    return DeviceArena(kind, n, n_games, sims, cpuct, **kw)


def rules_eval(kind, n, fl_map, states):
    lib = _lib.lib()
    B, A = states.shape[0], action_size(kind, n)
    dev = torch.device("cuda")
    s = torch.as_tensor(states).to(dev)
    valids = torch.zeros(B, dtype=torch.int32, device=dev)
    ended = torch.zeros(B, dtype=torch.float64, device=dev)
    etag = torch.zeros(B, dtype=torch.int8, device=dev)
    nxt = torch.zeros(B, A, 2, dtype=torch.int64, device=dev)
    _lib.check(lib.azg_rules_eval(GAME_KINDS[kind], n, fl_map, _lib.ptr(s), B, _lib.ptr(valids), _lib.ptr(ended),
                                  _lib.ptr(etag), _lib.ptr(nxt), _lib.stream()))
    return valids.cpu().numpy().astype(np.uint32), ended.cpu().numpy(), etag.cpu().numpy(), nxt.cpu().numpy()


@pytest.mark.parametrize("tag", ["c4_7_gnn", "c4_7_std", "c4_7_wide", "c4_5_gnn", "ttt_3_gnn", "ttt_4_std"])
def test_golden_episode(tag):
    cases.case_golden_episode(make_arena, tag)


def test_known_answer():
    cases.case_known_answer(make_arena)


@pytest.mark.parametrize("tag", ["c4_7", "ttt_4"])
def test_lockstep_games(tag):
    cases.case_lockstep_games(make_arena, tag)


@pytest.mark.parametrize("n", [4, 8])
def test_frozenlake(n):
    cases.case_frozenlake(make_arena, n)


@pytest.mark.parametrize("tag", ["c4_7", "c4_5", "c4_4", "ttt_3", "ttt_4", "fl_4", "fl_8"])
def test_rules(tag):
    cases.case_rules(rules_eval, tag)


def test_capacity_overflow():
    cases.case_capacity_overflow(make_arena)
