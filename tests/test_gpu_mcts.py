"""Tree-search parity: the GPU lock-step PV-MCTS against the unmodified reference pv_mcts_policy
(fixtures in tests/golden/mcts_golden.json, generated with deterministic evaluators so that priors
and values are bit-identical on both sides).  Visit counts must match exactly."""
import numpy as np
import pytest
import torch

from alphaquoridorgnn_b200 import game_logic as gl
from alphaquoridorgnn_b200 import pv_mcts
from alphaquoridorgnn_b200.pv_network_gnn import GNNNetwork

pytestmark = pytest.mark.gpu


def hash_evaluator(packed):
    """Same function as tests/golden/make_golden.py:hash_evaluator, in torch integer ops."""
    rows, plies = gl.unpack_rows(packed)
    w = (torch.arange(68, device=rows.device, dtype=torch.int64) + 1) * 7919
    key = ((rows.to(torch.int64) * w).sum(1) + plies.to(torch.int64) * 104729) % (2 ** 31)
    mask, pawn = gl.legal_mask_batch(packed)
    dense = gl.mask_to_dense(mask)
    a = torch.arange(209, device=rows.device, dtype=torch.int64)
    raw = (key[:, None] + a[None, :] * 40503) % 1009 + 1
    raw = torch.where(dense, raw, torch.zeros_like(raw))
    pri = raw.to(torch.float32) / raw.sum(1).to(torch.float32)[:, None]
    val = ((key % 2001) - 1000).to(torch.float32) / torch.tensor(1000.0, dtype=torch.float32, device=rows.device)
    return {"priors": pri.contiguous(), "value": val.contiguous(), "mask": mask, "pawn": pawn}


def uniform_evaluator(packed):
    mask, pawn = gl.legal_mask_batch(packed)
    dense = gl.mask_to_dense(mask)
    n = dense.sum(1).to(torch.float64)
    p = (1.0 / n).to(torch.float32)  # np.full(len, 1.0 / len, dtype=np.float32)
    pri = torch.where(dense, p[:, None].expand(-1, 209), torch.zeros((), device=packed.device))
    return {"priors": pri.contiguous(), "value": torch.zeros(packed.shape[0], device=packed.device), "mask": mask, "pawn": pawn}


EVAL = {"hash": hash_evaluator, "uniform": uniform_evaluator}


def test_visit_counts_match_reference(mcts_golden):
    for sims in (50, 200):
        for name in ("hash", "uniform"):
            cases = [r for r in mcts_golden["roots"] if r["sims"] == sims and r["evaluator"] == name]
            rows = np.array([c["row"] for c in cases], np.uint8)
            plies = np.array([c["plies"] for c in cases], np.int16)
            packed = gl.pack_rows(rows, plies)
            counts, actions, n = pv_mcts.BatchedMCTS(EVAL[name], sims).search(packed)
            for i, c in enumerate(cases):
                k = int(n[i])
                assert actions[i, :k].tolist() == c["legal_actions"], (name, sims, i)
                assert counts[i, :k].tolist() == c["visit_counts"], (name, sims, i)
                assert int(counts[i].sum()) == sims - 1


def test_batching_across_games_does_not_change_a_search(mcts_golden):
    cases = [r for r in mcts_golden["roots"] if r["sims"] == 50 and r["evaluator"] == "hash"]
    rows = np.array([c["row"] for c in cases], np.uint8)
    plies = np.array([c["plies"] for c in cases], np.int16)
    packed = gl.pack_rows(rows, plies)
    all_counts, _, _ = pv_mcts.BatchedMCTS(hash_evaluator, 50).search(packed)
    one_counts, _, _ = pv_mcts.BatchedMCTS(hash_evaluator, 50).search(packed[3:4])
    assert torch.equal(all_counts[3], one_counts[0])


def test_greedy_game_matches_reference(mcts_golden):
    g = mcts_golden["greedy_game"]
    old = pv_mcts.PV_EVALUATE_COUNT
    pv_mcts.PV_EVALUATE_COUNT = g["sims"]
    try:
        state = gl.State()
        seq = []
        for _ in range(len(g["actions"])):
            if state.is_done():
                break
            pol = pv_mcts.pv_mcts_policy(hash_evaluator, state, 0)
            a = state.legal_actions()[int(np.argmax(pol))]
            seq.append(int(a))
            state = state.next(a)
        assert seq == g["actions"]
        assert [*state.player, *state.enemy, *state.walls] == g["final_row"] and state.plies_played == g["final_plies"]
    finally:
        pv_mcts.PV_EVALUATE_COUNT = old


def test_policy_api_shapes_with_the_real_network():
    torch.manual_seed(0)
    net = GNNNetwork().cuda().eval()
    s = gl.State()
    pol = pv_mcts.pv_mcts_policy(net, s, 1.0)
    assert isinstance(pol, list) and len(pol) == 131 and abs(sum(pol) - 1.0) < 1e-12
    assert pv_mcts.boltzman([1, 2, 1], 1.0) == [0.25, 0.5, 0.25]
    act = pv_mcts.pv_mcts_action(net, 0)(s)
    assert act in s.legal_actions()


def test_cuda_graph_replay_equals_eager_search(mcts_golden):
    """The network path replays one captured simulation step; it must give the same tree as the eager loop."""
    torch.manual_seed(1)
    net = GNNNetwork().cuda().eval()
    cases = [r for r in mcts_golden["roots"] if r["sims"] == 50 and r["evaluator"] == "hash"]
    packed = gl.pack_rows(np.array([c["row"] for c in cases], np.uint8), np.array([c["plies"] for c in cases], np.int16))
    for prec in ("fp32", "bf16"):
        net.precision = prec
        graph = pv_mcts.BatchedMCTS(net, 40, use_graph=True)
        eager = pv_mcts.BatchedMCTS(net, 40, use_graph=False)
        c1, a1, n1 = graph.search(packed)
        c2, a2, n2 = eager.search(packed)
        c3, _, _ = graph.search(packed)  # second call replays the cached graph from simulation 1
        assert graph.use_graph and len(graph._graphs) == 1
        assert torch.equal(c1, c2) and torch.equal(a1, a2) and torch.equal(n1, n2) and torch.equal(c1, c3)
        assert int(c1.sum(1).min()) == 39


def test_search_on_a_state_without_legal_actions_is_defined(ka):
    """KA13-like position with no walls in hand: legal_actions() == [] (the reference would spin / crash at
    np.argmax([])).  Here the root is re-evaluated every simulation and has no children."""
    v = ka["KA13"]
    row = list(v["row"])
    row[1] = 0  # no walls in hand -> no wall actions either
    packed = gl.pack_rows(np.array([row], np.uint8), np.array([v["plies"]], np.int16))
    actions, n = gl.legal_actions_batch(packed)
    assert int(n[0]) == 0
    counts, acts, nc = pv_mcts.BatchedMCTS(uniform_evaluator, 8).search(packed)
    assert int(nc[0]) == 0 and int(counts.sum()) == 0 and int((acts >= 0).sum()) == 0
