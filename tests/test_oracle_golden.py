"""Pins the CPU oracle (oracle/) to fixtures produced by the unmodified reference."""
import numpy as np
import pytest

from oracle import quoridor_oracle as qo


def _row(v):
    return np.array(v, dtype=np.uint8)


def test_c_oracle_trajectories_exact(traj):
    out = qo.legal_actions_batch(traj["rows"], traj["plies"])
    assert np.array_equal(out["n"], traj["nact"])
    assert np.array_equal(out["mask"], traj["mask"])
    assert np.array_equal(out["pawn"], traj["pawn"])


def test_c_oracle_ordered_list_matches_mask_and_pawn(traj):
    out = qo.legal_actions_batch(traj["rows"][:4000], traj["plies"][:4000])
    for i in range(0, 4000, 37):
        n = int(out["n"][i])
        la = out["actions"][i, :n].tolist()
        k = int(traj["pawn"][i, 0])
        assert la[:k] == traj["pawn"][i, 1:1 + k].tolist()
        walls = [a for s in range(64) for a in (81 + s, 145 + s)
                 if (int(traj["mask"][i, a >> 5]) >> (a & 31)) & 1]
        assert la[k:] == walls
        assert (out["actions"][i, n:] == -1).all()


def test_c_oracle_next_follows_reference_games(traj):
    rows, plies, game, action = traj["rows"], traj["plies"], traj["game"], traj["action"]
    idx = np.nonzero((game[:-1] >= 0) & (game[:-1] == game[1:]) & (action[:-1] >= 0))[0]
    nxt, npl, flags = qo.next_batch(rows[idx], plies[idx], action[idx])
    assert np.array_equal(nxt, rows[idx + 1])
    assert np.array_equal(npl, plies[idx + 1])
    lose = (rows[idx + 1][:, 2] // 9) == 0
    assert np.array_equal((flags & 1).astype(bool), lose)
    assert np.array_equal(((flags >> 1) & 1).astype(bool), plies[idx + 1] >= 116)


def test_python_oracle_trajectories_subset(traj):
    sel = np.linspace(0, traj["rows"].shape[0] - 1, 700).astype(int)
    ref = qo.legal_actions_batch(traj["rows"][sel], traj["plies"][sel])
    for k, i in enumerate(sel):
        s = qo.PyOracleState.from_row68(traj["rows"][i], traj["plies"][i])
        n = int(ref["n"][k])
        assert s.legal_actions() == ref["actions"][k, :n].tolist()


def test_g1_and_ka_vectors(ka):
    g1 = ka["G1"]
    assert g1["legal_actions_wall"] == []
    assert qo.legal_actions_wall(_row(g1["row"]), g1["wall_pos"]) == g1["legal_actions_wall"]
    for name in ("G1", "KA1", "KA2", "KA3", "KA4", "KA5", "KA6", "KA13"):
        v = ka[name]
        out = qo.legal_actions_batch(_row(v["row"])[None], np.array([v["plies"]], np.int16))
        n = int(out["n"][0])
        assert out["actions"][0, :n].tolist() == v["legal_actions"], name
        s = qo.PyOracleState.from_row68(v["row"], v["plies"])
        assert s.legal_actions() == v["legal_actions"], name
        if "legal_actions_pos" in v:
            pos = v.get("pos", v["row"][0])
            assert qo.legal_actions_pos(_row(v["row"]), pos) == v["legal_actions_pos"], name
            assert s.legal_actions_pos(pos) == v["legal_actions_pos"], name
    # survey's hand-recorded values
    assert len(ka["KA1"]["legal_actions"]) == 131 and ka["KA1"]["legal_actions"][:7] == [67, 75, 77, 81, 145, 82, 146]
    assert ka["KA2"]["legal_actions_pos"] == [22, 49, 39, 41]
    assert ka["KA3"]["legal_actions_pos"] == [30, 32, 49, 39, 41]
    assert ka["KA4"]["legal_actions_pos"] == [3, 5, 22, 12, 14]
    assert ka["KA5"]["legal_actions_pos"] == [31, 49, 39, 42]
    assert ka["KA6"]["legal_actions_pos"] == [31, 49, 39, 32, 50]
    assert ka["KA13"]["legal_actions_pos"] == [] and len(ka["KA13"]["legal_actions"]) == 122


def test_ka_next_and_terminal(ka):
    for name in ("KA7", "KA8", "KA9"):
        v = ka[name]
        nxt, npl, _ = qo.next_batch(_row(v["row"])[None], np.array([v["plies"]]), np.array([v["action"]]))
        assert nxt[0].tolist() == v["next_row"] and int(npl[0]) == v["next_plies"], name
        s = qo.PyOracleState.from_row68(v["row"], v["plies"]).next(v["action"])
        assert s.row68().tolist() == v["next_row"] and s.plies_played == v["next_plies"]
    for row, want in ka["KA12"]["is_lose"]:
        assert qo.PyOracleState.from_row68(row).is_lose() == want
    for plies, want in ka["KA12"]["is_draw"]:
        assert qo.PyOracleState(plies_played=plies).is_draw() == want


def test_ka10_wall_rules(ka):
    v = ka["KA10"]
    for p, want in v["legal_actions_wall"].items():
        assert qo.legal_actions_wall(_row(v["row"]), int(p)) == want
        assert qo.PyOracleState.from_row68(v["row"]).legal_actions_wall(int(p)) == want


def test_planes_and_open_masks(ka, traj, graph_golden):
    v = ka["KA11"]
    planes = qo.planes_batch(_row(v["row"])[None])[0]
    assert np.array_equal(planes, np.array(v["pieces_array"], dtype=np.float32))
    idx = graph_golden["index"]
    assert np.array_equal(qo.planes_batch(traj["rows"][idx]), graph_golden["planes"])
    assert np.array_equal(qo.open_mask_batch(traj["rows"][idx]), graph_golden["open"])


def test_open_mask_symmetric_and_edge_count(traj):
    rows = traj["rows"][traj["game"] >= 0][::50]
    opn = qo.open_mask_batch(rows).reshape(-1, 9, 9)
    up, down, left, right = [(opn >> d) & 1 for d in range(4)]
    assert np.array_equal(up[:, 1:, :], down[:, :-1, :]) and not up[:, 0].any() and not down[:, 8].any()
    assert np.array_equal(left[:, :, 1:], right[:, :, :-1]) and not left[:, :, 0].any() and not right[:, :, 8].any()
    nwalls = (rows[:, 4:] != 0).sum(axis=1)
    edges = np.unpackbits(opn.reshape(len(rows), -1), axis=1).sum(axis=1)
    assert np.array_equal(edges, 288 - 4 * nwalls)  # legal positions: every wall severs 4 directed edges


def test_mcts_port_matches_reference_visit_counts(mcts_golden):
    """oracle/mcts_oracle.py (the CPU baseline of the MCTS metric) against the visit counts of the UNMODIFIED reference
    pv_mcts_policy (tests/golden/make_golden.py: deterministic hash / uniform evaluators)."""
    from oracle import mcts_oracle

    def hash_predict(state):
        r = [int(v) for v in state.row]
        key = (sum(v * (i + 1) * 7919 for i, v in enumerate(r)) + state.plies_played * 104729) % (2 ** 31)
        la = state.legal_actions()
        raw = np.array([(key + a * 40503) % 1009 + 1 for a in la], dtype=np.int64)
        return raw.astype(np.float32) / np.float32(raw.sum()), float(np.float32((key % 2001) - 1000) / np.float32(1000))

    def uniform_predict(state):
        la = state.legal_actions()
        return np.full(len(la), 1.0 / len(la), dtype=np.float32), 0.0

    checked = 0
    for case in mcts_golden["roots"]:
        if case["sims"] != 50 and checked >= 20:   # the 200-simulation cases: a few are enough for the CPU suite's time budget
            continue
        st = mcts_oracle.COracleState(np.array(case["row"], np.uint8), case["plies"])
        assert st.legal_actions() == case["legal_actions"]
        counts = mcts_oracle.pv_mcts_scores(hash_predict if case["evaluator"] == "hash" else uniform_predict, st, case["sims"])
        assert counts == case["visit_counts"], (case["evaluator"], case["sims"])
        checked += 1
    assert checked >= 16
