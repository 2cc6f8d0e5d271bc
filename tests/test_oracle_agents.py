"""Pins the oracle's restatement of agents.py (heuristic_eval, alpha_beta_action) to fixtures produced by the
unmodified reference (tests/golden/make_agents_golden.py)."""
import numpy as np

from oracle import quoridor_oracle as qo


def test_heuristic_eval_exact(traj, agents_golden):
    idx = np.array(agents_golden["heuristic"]["index"])
    dist, heur = qo.heuristic_batch(traj["rows"][idx])
    assert np.array_equal(heur, np.array(agents_golden["heuristic"]["value"]))  # float64, bit for bit
    assert np.array_equal(heur, (dist[:, 1].astype(np.int64) - dist[:, 0]) / 48)


def test_alpha_beta_action_exact(traj, agents_golden):
    assert len(agents_golden["alpha_beta"]) >= 200
    for case in agents_golden["alpha_beta"]:
        i = case["index"]
        a, _ = qo.alpha_beta_action(traj["rows"][i], traj["plies"][i], max_depth=case["max_depth"])
        assert a == case["action"], case
