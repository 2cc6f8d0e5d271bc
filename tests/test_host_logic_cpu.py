"""Host-side logic of the drivers that needs no GPU: visit counts -> search policy (pv_mcts.py:88-95, 106-109), the searcher cache
kept on a network object, first_player_value (self_play.py:22-27) and first_player_point (evaluate_network.py:18-22)."""
import copy
import pickle

import numpy as np
import torch

from alphaquoridorgnn_b200 import evaluate_network, pv_mcts, self_play
from oracle import quoridor_oracle as qo


def test_policy_from_counts_is_the_reference_boltzmann_distribution():
    rng = np.random.default_rng(0)
    counts = torch.from_numpy(rng.integers(0, 50, (7, 136)).astype(np.int32))
    counts[:, 100:] = 0                                   # beyond the root's children
    counts[3, :100] = 0
    counts[3, 17] = counts[3, 40] = 9                     # a tie: np.argmax takes the first
    for temperature in (1.0, 0.5, 2.0):
        pol = pv_mcts.policy_from_counts(counts, temperature).numpy()
        for g in range(7):
            ref = pv_mcts.boltzman([float(c) for c in counts[g, :100].tolist()], temperature)   # the reference's own expression
            assert np.allclose(pol[g, :100], ref, rtol=1e-14, atol=0) and not pol[g, 100:].any()
    greedy = pv_mcts.policy_from_counts(counts, 0).numpy()
    assert np.array_equal(greedy.argmax(axis=1), counts.numpy().argmax(axis=1)) and np.array_equal(greedy.sum(axis=1), np.ones(7))
    assert greedy[3, 17] == 1.0 and greedy[3, 40] == 0.0


def test_searcher_cache_is_not_copied_or_pickled_with_its_network():
    cache = pv_mcts._SearcherCache()
    cache[(200, "cuda:0")] = object()

    class Net:
        pass

    net = Net()
    net._searchers = cache
    twin = copy.deepcopy(net)
    assert isinstance(twin._searchers, pv_mcts._SearcherCache) and len(twin._searchers) == 0 and len(net._searchers) == 1
    assert len(pickle.loads(pickle.dumps(cache))) == 0


def test_game_result_conventions():
    lost_first = qo.PyOracleState(player=[40, 5], enemy=[4, 5], walls=[0] * 64, plies_played=10)    # enemy on row 0, first player to move
    lost_second = qo.PyOracleState(player=[40, 5], enemy=[4, 5], walls=[0] * 64, plies_played=11)
    draw = qo.PyOracleState(player=[40, 5], enemy=[40, 5], walls=[0] * 64, plies_played=116)
    assert lost_first.is_lose() and lost_first.is_first_player() and not draw.is_lose() and draw.is_draw()
    assert self_play.first_player_value(lost_first) == -1 and self_play.first_player_value(lost_second) == 1
    assert self_play.first_player_value(draw) == 0
    assert evaluate_network.first_player_point(lost_first) == 0 and evaluate_network.first_player_point(lost_second) == 1
    assert evaluate_network.first_player_point(draw) == 0.5
