"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py [--traj-games 420] [--synthetic 8000]

Outputs (committed):
  ka_vectors.json      G1 + KA1..KA13, inputs and reference outputs
  trajectories.npz     positions harvested from seeded random games played with the
                       reference's own legal_actions()/next(), with the reference answer for
                       every position; plus synthetic (not necessarily reachable) positions
  mcts_golden.json     pv_mcts_policy visit counts under a deterministic hash evaluator, and a
                       greedy MCTS game (action sequence)

The position generator follows SURVEY.md section 8d: game g uses random.Random(seed*2**32+g);
at each ply with probability 0.5 a uniformly random legal wall action (if any), otherwise a
uniformly random legal pawn action.
"""
import argparse
import json
import os
import random
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402

N = 9
NSQ = 81
NSLOT = 64


def row68(state):
    p, e, w = state.to_array()
    return [p[0], p[1], e[0], e[1]] + list(w)


def encode_actions(la):
    mask = [0] * 8
    for a in la:
        mask[a >> 5] |= 1 << (a & 31)
    pawn = [0] * 8
    k = 0
    while k < len(la) and la[k] < NSQ:
        pawn[1 + k] = la[k]
        k += 1
    for j in range(1 + k, 6):
        pawn[j] = 0xFF
    pawn[0] = k
    # the ordered list must be recoverable from (pawn order, mask): walls ascend by slot, H before V
    rebuilt = list(la[:k])
    for s in range(NSLOT):
        for a in (NSQ + s, NSQ + NSLOT + s):
            if (mask[a >> 5] >> (a & 31)) & 1:
                rebuilt.append(a)
    assert rebuilt == list(la), (rebuilt, la)
    return mask, pawn


def play_game(args):
    seed, g = args
    gl, _ = load_reference()
    rng = random.Random(seed * 2 ** 32 + g)
    state = gl.State()
    out = []
    while not state.is_done():
        la = state.legal_actions()
        mask, pawn = encode_actions(la)
        if not la:
            out.append((row68(state), state.plies_played, g, -1, len(la), mask, pawn))
            break
        walls = [a for a in la if a >= NSQ]
        pawns = [a for a in la if a < NSQ]
        if walls and (rng.random() < 0.5 or not pawns):
            a = rng.choice(walls)
        else:
            a = rng.choice(pawns)
        out.append((row68(state), state.plies_played, g, a, len(la), mask, pawn))
        state = state.next(a)
    # terminal state too (legal_actions is still defined on it)
    la = state.legal_actions()
    mask, pawn = encode_actions(la)
    out.append((row68(state), state.plies_played, g, -1, len(la), mask, pawn))
    return out


def synthetic_state(args):
    seed, i = args
    gl, _ = load_reference()
    rng = random.Random(seed * 2 ** 32 + 10 ** 6 + i)
    walls = [0] * NSLOT
    target = rng.randint(0, 20)
    st = gl.State(board_size=N, player=[40, 1], enemy=[40, 1], walls=walls)
    placed = 0
    for _ in range(400):
        if placed >= target:
            break
        pos, o = rng.randrange(NSLOT), rng.choice((1, 2))
        # placement rule only (game_logic.py:199-223), paths are NOT checked: the point of this
        # set is to hit blocked / cornered / gate-sensitive positions
        if walls[pos] != 0:
            continue
        x, y = divmod(pos, 8)
        if o == 1 and ((y > 0 and walls[pos - 1] == 1) or (y < 7 and walls[pos + 1] == 1)):
            continue
        if o == 2 and ((x > 0 and walls[pos - 8] == 2) or (x < 7 and walls[pos + 8] == 2)):
            continue
        walls[pos] = o
        placed += 1
    ppos = rng.randrange(9, NSQ) if rng.random() < 0.9 else rng.randrange(NSQ)
    while True:
        epos = rng.randrange(9, NSQ) if rng.random() < 0.9 else rng.randrange(NSQ)
        if rng.random() < 0.35:  # force adjacency often so the jump rules are exercised
            x, y = divmod(ppos, N)
            dx, dy = rng.choice(((-1, 0), (1, 0), (0, -1), (0, 1)))
            if 0 <= x + dx < N and 0 <= y + dy < N:
                epos = 80 - ((x + dx) * N + y + dy)
        if 80 - epos != ppos:
            break
    pw = rng.choice((0, 1, 1, 2, 5, 10))
    ew = rng.randint(0, 10)
    plies = rng.randint(0, 115)
    st = gl.State(board_size=N, player=[ppos, pw], enemy=[epos, ew], walls=walls, plies_played=plies)
    la = st.legal_actions()
    mask, pawn = encode_actions(la)
    return (row68(st), plies, -1, -1, len(la), mask, pawn)


def hash_evaluator(gl):
    """Deterministic stand-in for model.predict (BaseNetwork.py:36-40 contract): float32
    priors over legal actions (normalised) and a python-float value, both pure functions of
    the state.  tests/ re-implement the same function with torch integer ops."""

    class Fake:
        def predict(self, state, device):
            r = row68(state)
            key = (sum(int(v) * (i + 1) * 7919 for i, v in enumerate(r)) + state.plies_played * 104729) % (2 ** 31)
            la = state.legal_actions()
            raw = np.array([(key + a * 40503) % 1009 + 1 for a in la], dtype=np.int64)
            pri = raw.astype(np.float32) / np.float32(raw.sum())
            val = float(np.float32((key % 2001) - 1000) / np.float32(1000))
            return pri, val

    return Fake()


def make_ka(gl):
    S = gl.State
    ka = {}

    def st(player=None, enemy=None, walls=None, plies=0):
        w = [0] * NSLOT
        for k, v in (walls or {}).items():
            w[int(k)] = v
        if player is None:
            s = S(board_size=N, num_walls=10, walls=w, plies_played=plies)
        else:
            s = S(board_size=N, player=list(player), enemy=list(enemy), walls=w, plies_played=plies)
        return s

    # G1: test_legal_walls.py:1-21
    g1 = st(walls={24: 1, 27: 1, 32: 2, 36: 2, 37: 1, 41: 1, 42: 2, 43: 1})
    g1.player[0] = 40
    g1.enemy[0] = 32
    ka["G1"] = {"row": row68(g1), "plies": 0, "wall_pos": 26, "legal_actions_wall": g1.legal_actions_wall(26),
                "legal_actions": g1.legal_actions()}
    s = st()
    ka["KA1"] = {"row": row68(s), "plies": 0, "legal_actions": s.legal_actions()}
    for name, player, enemy, walls, pos in (
            ("KA2", [40, 10], [49, 10], {}, 40),
            ("KA3", [40, 10], [49, 10], {19: 1}, 40),
            ("KA4", [13, 10], [76, 10], {}, 13),
            ("KA5", [40, 10], [39, 10], {}, 40),
            ("KA6", [40, 10], [39, 10], {37: 2}, 40)):
        s = st(player, enemy, walls)
        ka[name] = {"row": row68(s), "plies": 0, "pos": pos, "legal_actions_pos": s.legal_actions_pos(pos),
                    "legal_actions": s.legal_actions()}
    for name, action in (("KA7", 67), ("KA8", 81), ("KA9", 81 + 64 + 9)):
        s = st()
        t = s.next(action)
        ka[name] = {"row": row68(s), "plies": 0, "action": action, "next_row": row68(t),
                    "next_plies": t.plies_played}
    s = st(walls={10: 1})
    ka["KA10"] = {"row": row68(s), "plies": 0,
                  "legal_actions_wall": {str(p): s.legal_actions_wall(p) for p in (9, 10, 11, 2, 18)}}
    s = st([76, 10], [70, 9], {10: 1})
    ka["KA11"] = {"row": row68(s), "plies": 0,
                  "pieces_array": np.array(s.pieces_array(), dtype=np.int64).reshape(6, N, N).tolist()}
    ka["KA12"] = {
        "is_lose": [[row68(st([40, 5], [5, 5])), True], [row68(st([40, 5], [9, 5])), False]],
        "is_draw": [[116, st(plies=116).is_draw()], [115, st(plies=115).is_draw()]],
    }
    s = st([72, 5], [17, 5], {56: 2, 48: 1}, plies=10)
    ka["KA13"] = {"row": row68(s), "plies": 10, "legal_actions_pos": s.legal_actions_pos(72),
                  "legal_actions": s.legal_actions()}
    return ka


def make_mcts(gl, mc, roots):
    out = {"roots": [], "greedy_game": None}
    fake = hash_evaluator(gl)

    class Uniform:
        def predict(self, state, device):
            la = state.legal_actions()
            return np.full(len(la), 1.0 / len(la), dtype=np.float32), 0.0

    for sims in (50, 200):
        mc.PV_EVALUATE_COUNT = sims
        for name, model in (("hash", fake), ("uniform", Uniform())):
            for (row, plies) in roots:
                s = gl.State(board_size=N, player=row[0:2], enemy=row[2:4], walls=row[4:], plies_played=plies)
                pol = mc.pv_mcts_policy(model, s, 1.0, "cpu")
                counts = [int(round(float(p) * (sims - 1))) for p in pol]
                assert sum(counts) == sims - 1
                out["roots"].append({"row": row, "plies": plies, "sims": sims, "evaluator": name,
                                     "legal_actions": s.legal_actions(), "visit_counts": counts})
    # greedy game: temperature 0, action = first most-visited child (pv_mcts.py:88-92)
    mc.PV_EVALUATE_COUNT = 30
    s = gl.State()
    seq = []
    for _ in range(60):
        if s.is_done():
            break
        pol = mc.pv_mcts_policy(fake, s, 0, "cpu")
        a = s.legal_actions()[int(np.argmax(pol))]
        seq.append(int(a))
        s = s.next(a)
    out["greedy_game"] = {"sims": 30, "actions": seq, "final_row": row68(s), "final_plies": s.plies_played}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--traj-games", type=int, default=420)
    ap.add_argument("--synthetic", type=int, default=8000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--skip-mcts", action="store_true")
    args = ap.parse_args()
    gl, mc = load_reference()

    ka = make_ka(gl)
    with open(os.path.join(HERE, "ka_vectors.json"), "w") as f:
        json.dump(ka, f)
    print("ka_vectors.json written")

    with Pool(8) as pool:
        games = pool.map(play_game, [(args.seed, g) for g in range(args.traj_games)], chunksize=4)
        synth = pool.map(synthetic_state, [(args.seed, i) for i in range(args.synthetic)], chunksize=64)
    recs = [r for g in games for r in g] + synth
    rows = np.array([r[0] for r in recs], dtype=np.uint8)
    plies = np.array([r[1] for r in recs], dtype=np.int16)
    game = np.array([r[2] for r in recs], dtype=np.int32)
    action = np.array([r[3] for r in recs], dtype=np.int16)
    nact = np.array([r[4] for r in recs], dtype=np.int16)
    mask = np.array([r[5] for r in recs], dtype=np.uint32)
    pawn = np.array([r[6] for r in recs], dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "trajectories.npz"), rows=rows, plies=plies, game=game,
                        action=action, nact=nact, mask=mask, pawn=pawn)
    print("trajectories.npz:", rows.shape[0], "positions,", int((game >= 0).sum()), "from games;",
          "walls-in-hand:", int((rows[:, 1] > 0).sum()), " zero-legal:", int((nact == 0).sum()))

    # planes + open-direction masks straight from the reference for a subset
    sub = np.linspace(0, rows.shape[0] - 1, 600).astype(np.int64)
    planes = np.zeros((len(sub), 6, N, N), np.float32)
    openm = np.zeros((len(sub), NSQ), np.uint8)
    for k, i in enumerate(sub):
        r = rows[i].tolist()
        s = gl.State(board_size=N, player=r[0:2], enemy=r[2:4], walls=r[4:], plies_played=int(plies[i]))
        planes[k] = np.array(s.pieces_array(), dtype=np.float32).reshape(6, N, N)
        for v in range(NSQ):
            # edge oracle: park the enemy far from v so no jump rule fires
            far = 0 if (v // N) >= 4 else 80
            probe = gl.State(board_size=N, player=[v, 0], enemy=[80 - far, 0], walls=r[4:])
            m = 0
            for q in probe.legal_actions_pos(v):
                d = {-9: 0, 9: 1, -1: 2, 1: 3}[q - v]
                m |= 1 << d
            openm[k, v] = m
    np.savez_compressed(os.path.join(HERE, "graph_golden.npz"), index=sub, planes=planes, open=openm)
    print("graph_golden.npz written")

    if not args.skip_mcts:
        cand = [i for i in range(rows.shape[0]) if game[i] >= 0 and action[i] >= 0 and nact[i] > 0 and rows[i, 2] // N != 0]
        rng = random.Random(7)
        picks = [0] + rng.sample([i for i in cand if rows[i, 1] > 0], 4) + rng.sample([i for i in cand if rows[i, 1] == 0], 2)
        roots = [(rows[i].tolist(), int(plies[i])) for i in picks]
        m = make_mcts(gl, mc, roots)
        with open(os.path.join(HERE, "mcts_golden.json"), "w") as f:
            json.dump(m, f)
        print("mcts_golden.json written:", len(m["roots"]), "roots; greedy game", len(m["greedy_game"]["actions"]), "plies")


if __name__ == "__main__":
    main()
