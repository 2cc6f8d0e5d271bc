"""Generate tests/golden/agents_golden.json from the UNMODIFIED reference agents.py.

Run in the build container only (needs /root/reference):  python tests/golden/make_agents_golden.py

Positions are taken from trajectories.npz (reachable positions harvested from seeded reference games).
  heuristic : agents.heuristic_eval(state) (float64) for 4,000 positions
  alpha_beta: agents.alpha_beta_action(state, max_depth) for positions chosen so that CPython finishes:
              max_depth 0 (200 positions), 1 (60 positions), 2 (12 positions whose mover has no walls left and 4
              with walls in hand)
"""
import importlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402


def main():
    gl, _ = load_reference()
    sys.modules.pop("agents", None)
    agents = importlib.import_module("agents")
    assert agents.MAX_DIST_FROM_GOAL == 48
    d = np.load(os.path.join(HERE, "trajectories.npz"))
    rows, plies, game = d["rows"], d["plies"], d["game"]
    reachable = np.nonzero(game >= 0)[0] if (game < 0).any() else np.arange(len(rows))
    rng = np.random.RandomState(7)

    def state_of(i):
        r = rows[i]
        return gl.State(player=[int(r[0]), int(r[1])], enemy=[int(r[2]), int(r[3])], walls=[int(x) for x in r[4:]],
                        plies_played=int(plies[i]))

    out = {"heuristic": {"index": [], "value": []}, "alpha_beta": []}
    t0 = time.time()
    for i in sorted(rng.choice(reachable, 4000, replace=False).tolist()):
        s = state_of(i)
        if s.is_done():
            continue
        before = s.to_array()
        v = agents.heuristic_eval(s)
        assert s.to_array() == before
        out["heuristic"]["index"].append(int(i))
        out["heuristic"]["value"].append(float(v))
    print("heuristic", len(out["heuristic"]["index"]), time.time() - t0, flush=True)

    live = [i for i in reachable.tolist() if not state_of(i).is_done()]
    live = np.array(live)
    nowalls = live[rows[live, 1] == 0]
    withwalls = live[rows[live, 1] > 0]
    picks = ([(int(i), 0) for i in rng.choice(live, 200, replace=False)] +
             [(int(i), 1) for i in rng.choice(live, 60, replace=False)] +
             [(int(i), 2) for i in rng.choice(nowalls, 12, replace=False)] +
             [(int(i), 2) for i in rng.choice(withwalls, 4, replace=False)])
    for i, depth in picks:
        t0 = time.time()
        a = agents.alpha_beta_action(state_of(i), max_depth=depth)
        out["alpha_beta"].append({"index": i, "max_depth": depth, "action": int(a)})
        print(i, depth, a, f"{time.time() - t0:.1f}s", flush=True)
    with open(os.path.join(HERE, "agents_golden.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
