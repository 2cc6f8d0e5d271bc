"""Golden fixture for the arena (evaluate_network.py:18-22, 25-45, 66-73 of the UNMODIFIED reference).

Run in the build container only (needs /root/reference):   python tests/golden/make_arena_golden.py

Two deterministic stand-ins for ``model.predict`` play the reference's own ``play(next_actions)`` through the reference's
``pv_mcts_action`` at temperature 0 (the most-visited child, so ``np.random.choice`` has a one-hot distribution and the games
are reproducible), in the colour order of the reference's match loop: game i even -> (model0, model1), odd -> reversed and
``1 - point``.  Output: tests/golden/arena_golden.json with, per pairing, every game's action sequence, the point the
reference's ``play`` returned and the total.  tests/test_gpu_drivers.py replays the pairings through the product's
``evaluate_network.play_matches`` with the same evaluators written in torch integer ops.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402
from make_golden import row68  # noqa: E402

SIMS = 16


def runner_weights(key, la, ppos, salt):
    """Integer prior weights of the 'runner' evaluator: forward pawn moves 60, sideways 6, backward 2, walls 1, plus a
    state-dependent jitter in 0..6 that breaks ties.  Pure integer arithmetic: reproducible bit for bit in torch."""
    raw = []
    for a in la:
        if a < 81:
            base = 60 if a // 9 < ppos // 9 else (6 if a // 9 == ppos // 9 else 2)
        else:
            base = 1
        raw.append(base * 8 + (key + a * 40503 + salt) % 7)
    return np.array(raw, dtype=np.int64)


def make_evaluator(kind, salt):
    class Fake:
        def predict(self, state, device):
            r = row68(state)
            key = (sum(int(v) * (i + 1) * 7919 for i, v in enumerate(r)) + state.plies_played * 104729) % (2 ** 31)
            la = state.legal_actions()
            if kind == "runner":
                raw = runner_weights(key, la, r[0], salt)
                val = 0.0
            else:  # "hash": the evaluator of make_golden.py with a salt
                raw = np.array([(key + a * 40503 + salt) % 1009 + 1 for a in la], dtype=np.int64)
                val = float(np.float32(((key + salt) % 2001) - 1000) / np.float32(1000))
            pri = raw.astype(np.float32) / np.float32(raw.sum())
            return pri, val

    return Fake()


PAIRINGS = [  # (name, (kind0, salt0), (kind1, salt1), games)
    ("runner_vs_hash", ("runner", 0), ("hash", 11), 4),
    ("runner_vs_runner", ("runner", 0), ("runner", 3), 5),
    ("hash_vs_hash", ("hash", 5), ("hash", 11), 2),
]


def main():
    gl, mc = load_reference()
    import importlib
    ev = importlib.import_module("evaluate_network")  # the reference's module (pv_network_cnn imports are stubbed by ref_loader)
    mc.PV_EVALUATE_COUNT = SIMS
    out = {"sims": SIMS, "temperature": 0, "pairings": []}
    for name, m0, m1, games in PAIRINGS:
        a0 = mc.pv_mcts_action(make_evaluator(*m0), 0, "cpu")
        a1 = mc.pv_mcts_action(make_evaluator(*m1), 0, "cpu")
        rec = {"name": name, "model0": list(m0), "model1": list(m1), "games": []}
        total = 0.0
        for i in range(games):  # evaluate_network.py:66-73
            log = []

            def wrap(f):
                def g(state):
                    a = int(f(state))
                    log.append(a)
                    return a
                return g

            nexts = (wrap(a0), wrap(a1))
            if i % 2 == 0:
                pt = ev.play(nexts)
                total += pt
            else:
                pt = ev.play(list(reversed(nexts)))
                total += 1 - pt
            rec["games"].append({"first_player_point": pt, "actions": log})
            print(name, "game", i, "plies", len(log), "first_player_point", pt, flush=True)
        rec["total_point_model0"] = total
        out["pairings"].append(rec)
    with open(os.path.join(HERE, "arena_golden.json"), "w") as f:
        json.dump(out, f)
    print("arena_golden.json written")


if __name__ == "__main__":
    main()
