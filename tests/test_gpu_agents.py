"""GPU parity of the agents row: shortest-path distances / heuristic_eval and the
level-synchronous negamax behind alpha_beta_action, bit-exact against the reference's golden output and against
the CPU oracle's literal fail-hard alpha-beta."""
import numpy as np
import pytest
import torch

from alphaquoridorgnn_b200 import agents
from alphaquoridorgnn_b200 import game_logic as gl
from alphaquoridorgnn_b200 import positions
from oracle import quoridor_oracle as qo

pytestmark = pytest.mark.gpu


def test_heuristic_eval_golden_and_oracle(traj, agents_golden):
    idx = np.array(agents_golden["heuristic"]["index"])
    packed = gl.pack_rows(traj["rows"][idx], traj["plies"][idx])
    out = agents.shortest_paths_batch(packed, want_leaf=True)
    assert np.array_equal(out["heuristic"].cpu().numpy(), np.array(agents_golden["heuristic"]["value"]))
    # every golden trajectory position (incl. synthetic, possibly walled-in ones: -1 = no path) against the oracle
    packed = gl.pack_rows(traj["rows"], traj["plies"])
    out = agents.shortest_paths_batch(packed, want_leaf=True)
    dist, heur = qo.heuristic_batch(traj["rows"])
    assert np.array_equal(out["dist"].cpu().numpy(), dist)
    assert np.array_equal(out["heuristic"].cpu().numpy(), heur)
    lose = (traj["rows"][:, 2] // 9) == 0
    draw = traj["plies"] >= 116
    want = np.where(lose, -48, np.where(draw, 0, dist[:, 1].astype(np.int32) - dist[:, 0]))
    assert np.array_equal(out["leaf48"].cpu().numpy(), want)


def test_state_wrappers_match_reference_values(traj, agents_golden):
    i = agents_golden["heuristic"]["index"][5]
    r = traj["rows"][i]
    s = gl.State(player=[int(r[0]), int(r[1])], enemy=[int(r[2]), int(r[3])], walls=[int(x) for x in r[4:]],
                 plies_played=int(traj["plies"][i]))
    assert agents.heuristic_eval(s) == agents_golden["heuristic"]["value"][5]
    assert agents.random_action(s) in s.legal_actions()
    case = next(c for c in agents_golden["alpha_beta"] if c["max_depth"] == 1)
    r = traj["rows"][case["index"]]
    s = gl.State(player=[int(r[0]), int(r[1])], enemy=[int(r[2]), int(r[3])], walls=[int(x) for x in r[4:]],
                 plies_played=int(traj["plies"][case["index"]]))
    assert agents.alpha_beta_action(s, max_depth=1) == case["action"]


def test_alpha_beta_action_golden(traj, agents_golden):
    for depth in (0, 1, 2):
        cases = [c for c in agents_golden["alpha_beta"] if c["max_depth"] == depth]
        idx = np.array([c["index"] for c in cases])
        packed = gl.pack_rows(traj["rows"][idx], traj["plies"][idx])
        got = []
        for k in range(0, len(idx), 4 if depth == 2 else 64):  # bounded level sizes
            got += agents.negamax_batch(packed[k:k + (4 if depth == 2 else 64)].contiguous(), max_depth=depth)["action"].tolist()
        assert got == [c["action"] for c in cases], depth


def test_negamax_scores_match_literal_alpha_beta_where_inside_the_window():
    """The oracle's literal alpha-beta reports the score seen for every root action; a score that raised alpha is
    exact, every other one is only an upper bound <= the running best.  Depth 1 on 96 GPU-generated positions."""
    packed = positions.random_positions(96, seed=11, games=8)
    rows, plies = [t.cpu().numpy() for t in gl.unpack_rows(packed)]
    out = agents.negamax_batch(packed, max_depth=1)
    off = out["offsets"].cpu().numpy()
    sc = out["scores48"].cpu().numpy()
    for b in range(96):
        a, scores = qo.alpha_beta_action(rows[b], plies[b], max_depth=1)
        assert int(out["action"][b]) == a
        mine = sc[off[b]:off[b + 1]] / 48.0
        ref = scores[:len(mine)]
        best = -np.inf
        for k in range(len(mine)):
            if ref[k] > best:
                assert mine[k] == ref[k]
                best = ref[k]
            else:
                assert mine[k] <= best
        assert int(out["value48"][b]) / 48.0 == best


def test_negamax_empty_and_terminal_children():
    assert agents.negamax_batch(torch.empty((0, 32), dtype=torch.uint8, device="cuda"), max_depth=1)["action"].numel() == 0
    # mover one step from the goal row: the winning pawn move must be chosen at any depth (child is_lose -> -1 -> score +1)
    rows = np.zeros((1, 68), np.uint8)
    rows[0, :4] = [13, 0, 40, 0]
    packed = gl.pack_rows(rows, np.array([30], np.int16))
    for depth in (0, 1, 2):
        out = agents.negamax_batch(packed, max_depth=depth)
        assert int(out["action"][0]) == 4 and int(out["value48"][0]) == 48


def test_shortest_paths_200k_generated_positions_against_oracle():
    """Row f4 at scale: 200,000 GPU-generated reachable positions (plies 0..~24 of 8,192 lock-step random games), both distances
    and the float64 heuristic against the CPU oracle's queue BFS on the rotated board -- bit-exact."""
    packed = positions.random_positions(200_000, seed=21, games=8192)
    rows, _ = [t.cpu().numpy() for t in gl.unpack_rows(packed)]
    out = agents.shortest_paths_batch(packed)
    dist, heur = qo.heuristic_batch(rows)
    assert np.array_equal(out["dist"].cpu().numpy(), dist)
    assert np.array_equal(out["heuristic"].cpu().numpy(), heur)
    # linearity property that needs no oracle: swapping the players' roles negates the heuristic (agents.py:43-52)
    r2 = rows.copy()
    r2[:, 0:2], r2[:, 2:4] = rows[:, 2:4], rows[:, 0:2]
    r2[:, 4:] = rows[:, 4:][:, ::-1]
    swapped = agents.heuristic_eval_batch(gl.pack_rows(r2, np.zeros(len(r2), np.int16)))
    assert np.array_equal(swapped.cpu().numpy(), -out["heuristic"].cpu().numpy())
