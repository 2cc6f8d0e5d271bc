"""GPU tests of the callers either side of the hot path: self-play history, training loop (flat-buffer
trainer vs torch autograd + torch.optim.Adam), arena evaluation and one tiny train cycle."""
import copy
import os
import pickle

import numpy as np
import pytest
import torch

from alphaquoridorgnn_b200 import game_logic as gl
from alphaquoridorgnn_b200 import evaluate_network, pv_mcts, self_play, train_network
from alphaquoridorgnn_b200.pv_network_gnn import GNNNetwork, create_network
from oracle import quoridor_oracle as qo

pytestmark = pytest.mark.gpu


def _net(seed=0):
    torch.manual_seed(seed)
    return GNNNetwork().cuda().eval()


def test_self_play_history_format_and_semantics():
    net = _net()
    history, info = self_play.play_batch(net, 6, sims=8, seed=1)
    assert len(history) == int(info["plies"].sum())
    # reference format: [[player, enemy, walls], policy[209], value]
    k = 0
    for g in range(6):
        n = int(info["plies"][g])
        game = history[k:k + n]
        k += n
        state = qo.PyOracleState()
        for i, (s, policy, value) in enumerate(game):
            assert s == state.to_array()                                   # trajectory is a legal game
            la = state.legal_actions()
            pol = np.array(policy)
            assert len(policy) == 209 and abs(pol.sum() - 1.0) < 1e-9
            assert set(np.nonzero(pol)[0].tolist()) <= set(la)            # mass only on legal actions
            assert value == game[0][2] * (1 if i % 2 == 0 else -1)        # alternating sign
            # the move actually played is not stored; recover it from the next state
            if i + 1 < n:
                nxt = game[i + 1][0]
                cand = [a for a in la if state.next(a).to_array() == nxt]
                assert cand, "next position is not a successor"
                state = state.next(cand[0])
        assert game[0][2] in (-1, 0, 1)
        lose, draw = info["flags"][g] & 1, info["flags"][g] & 2
        assert lose or draw
        if lose:   # the player to move in the final state lost: first player iff plies even
            assert game[0][2] == (-1 if n % 2 == 0 else 1)
        else:
            assert game[0][2] == 0 and n == 116


_M64 = (1 << 64) - 1


def _mix64(z):
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def _uniform(seed, game, ply):
    """The counter-based uniform aq_selfplay_advance documents: a hash of (seed, game id, ply) -> [0, 1)."""
    z = _mix64((seed + 0x9E3779B97F4A7C15 * (game + 1)) & _M64)
    z = _mix64((z + 0xD1B54A32D192ED03 * (ply + 1)) & _M64)
    return (z >> 11) / 9007199254740992.0


def _advance(states, counts, actions, n, game_id, temperature, seed, ply, num_games, dtype=torch.float64):
    from alphaquoridorgnn_b200 import _lib
    L, P = _lib.load(), _lib.ptr
    G = states.shape[0]
    dev = states.device
    out = dict(policy=torch.full((G, 209), -1, dtype=dtype, device=dev), chosen=torch.full((G,), -7, dtype=torch.int16, device=dev),
               nxt=torch.zeros_like(states), nxt_id=torch.full_like(game_id, -1), flags=torch.zeros(num_games, dtype=torch.uint8, device=dev),
               plies=torch.zeros(num_games, dtype=torch.int64, device=dev), alive=torch.zeros(1, dtype=torch.int32, device=dev))
    ws = torch.empty(L.aq_selfplay_ws_bytes(G), dtype=torch.uint8, device=dev)
    _lib.check(L.aq_selfplay_advance(P(states), P(counts), P(actions), P(n), P(game_id), G, float(temperature), seed, ply, P(out["policy"]),
                                     int(dtype == torch.float64), P(out["chosen"]), P(out["nxt"]), P(out["nxt_id"]), P(out["flags"]),
                                     P(out["plies"]), P(out["alive"]), P(ws), _lib.stream_ptr()), "aq_selfplay_advance")
    torch.cuda.synchronize()
    return out


def test_selfplay_advance_policy_move_and_compaction(traj):
    """aq_selfplay_advance against the reference's expressions evaluated on the host (self_play.py:47-60, pv_mcts.py:88-95):
    scores = counts / sum exactly, dense 209-wide record, the drawn action = inverse CDF of the documented uniform in
    legal_actions() order, next state = State.next, survivors in order, final flags / lengths of the games that ended."""
    rng = np.random.default_rng(3)
    # 1,500 > 1,024 exercises the second block of the compaction scan; 300 positions one ply before the draw so that games end
    live = np.nonzero((traj["plies"] < 116) & (traj["rows"][:, 2] // 9 != 0))[0]
    late = live[np.argsort(-traj["plies"][live], kind="stable")[:300]]
    pick = np.concatenate([late, rng.choice(np.setdiff1d(live, late), 1200, replace=False)])
    rng.shuffle(pick)
    rows, plies = traj["rows"][pick], traj["plies"][pick]
    G = len(rows)
    packed = gl.pack_rows(rows, plies, "cuda")
    legal = qo.legal_actions_batch(rows, plies)
    actions, nch = legal["actions"], legal["n"]
    counts = np.zeros((G, 136), dtype=np.int32)
    for g in range(G):
        k = int(nch[g])
        c = rng.integers(0, 40, k) * (rng.random(k) < 0.5)                 # many zero counts, as after a search
        if c.sum() == 0:
            c[rng.integers(k)] = 1
        counts[g, :k] = c
    game_id = torch.from_numpy(rng.permutation(5000)[:G].astype(np.int64)).cuda()
    seed, ply = 0xDEADBEEF12345, 57
    dev = lambda a: torch.from_numpy(a).cuda()
    out = _advance(packed, dev(counts), dev(actions), dev(nch), game_id, 1.0, seed, ply, 5000)
    pol = out["policy"].cpu().numpy()
    chosen = out["chosen"].cpu().numpy()
    gid = game_id.cpu().numpy()
    for g in range(G):
        k = int(nch[g])
        scores = counts[g, :k] / counts[g, :k].sum()                       # float64, exact for integer counts
        dense = np.zeros(209)
        dense[actions[g, :k]] = scores
        assert np.array_equal(pol[g], dense), g
        target = _uniform(seed, int(gid[g]), ply) * float(counts[g, :k].sum())
        cum = np.cumsum(counts[g, :k].astype(np.float64))
        hit = np.nonzero((counts[g, :k] > 0) & (target < cum))[0]
        assert chosen[g] == actions[g, hit[0]], g
    nrows_exp, nplies_exp, flags = qo.next_batch(rows, plies, chosen)
    live = flags == 0
    alive_rows, alive_ids = nrows_exp[live], gid[live]
    exp_flags, exp_plies = np.zeros(5000, np.uint8), np.zeros(5000, np.int64)
    exp_flags[gid[~live]], exp_plies[gid[~live]] = flags[~live], ply + 1
    k = int(out["alive"].item())
    assert 0 < k == len(alive_rows) < G                                    # the case has both survivors and finished games
    nrows, nplies = gl.unpack_rows(out["nxt"][:k].contiguous())
    assert np.array_equal(nrows.cpu().numpy(), alive_rows) and np.array_equal(nplies.cpu().numpy(), nplies_exp[live])
    assert np.array_equal(out["nxt_id"][:k].cpu().numpy(), alive_ids)
    assert np.array_equal(out["flags"].cpu().numpy(), exp_flags) and np.array_equal(out["plies"].cpu().numpy(), exp_plies)
    # float32 record: the same scores rounded once
    out32 = _advance(packed, dev(counts), dev(actions), dev(nch), game_id, 1.0, seed, ply, 5000, torch.float32)
    assert np.array_equal(out32["policy"].cpu().numpy(), pol.astype(np.float32)) and torch.equal(out32["chosen"], out["chosen"])
    # temperature 0: one-hot on the first maximum (np.argmax, pv_mcts.py:90-93)
    out0 = _advance(packed, dev(counts), dev(actions), dev(nch), game_id, 0.0, seed, ply, 5000)
    first_max = actions[np.arange(G), counts.argmax(axis=1)]
    assert np.array_equal(out0["chosen"].cpu().numpy(), first_max)
    p0 = out0["policy"].cpu().numpy()
    assert np.array_equal(p0.sum(axis=1), np.ones(G)) and np.array_equal(p0[np.arange(G), first_max], np.ones(G))
    # a game's draw depends on (seed, game, ply) only -- not on its row or on the other games of the batch
    sub = torch.arange(G - 1, -1, -3).cuda()
    outs = _advance(packed[sub].contiguous(), dev(counts)[sub].contiguous(), dev(actions)[sub].contiguous(), dev(nch)[sub].contiguous(),
                    game_id[sub].contiguous(), 1.0, seed, ply, 5000)
    assert torch.equal(outs["chosen"], out["chosen"][sub])
    # temperature 2: counts ** 0.5 normalised
    out2 = _advance(packed, dev(counts), dev(actions), dev(nch), game_id, 2.0, seed, ply, 5000)
    x = np.sqrt(counts.astype(np.float64))
    exp = np.zeros((G, 209))
    for g in range(G):
        exp[g, actions[g, :nch[g]]] = x[g, :nch[g]] / x[g, :nch[g]].sum()
    assert np.abs(out2["policy"].cpu().numpy() - exp).max() < 1e-14


def test_selfplay_advance_draws_follow_the_search_policy():
    """20,000 games with the same root and the same visit counts: the empirical frequencies of the drawn actions match
    counts / sum (np.random.choice(legal_actions, p=scores), self_play.py:57)."""
    G = 20000
    st = qo.PyOracleState()
    la = st.legal_actions()
    rows, plies = gl.rows_from_states([st])
    packed = gl.pack_rows(rows, plies, "cuda").expand(G, -1).contiguous()
    c = np.zeros(136, dtype=np.int32)
    c[:len(la)] = np.random.default_rng(1).integers(0, 30, len(la))
    a = np.full(136, -1, dtype=np.int16)
    a[:len(la)] = la
    rep = lambda v: torch.from_numpy(np.broadcast_to(v, (G,) + v.shape).copy()).cuda()
    out = _advance(packed, rep(c), rep(a), torch.full((G,), len(la), dtype=torch.int16).cuda(), torch.arange(G).cuda(), 1.0, 99, 0, G)
    freq = np.bincount(out["chosen"].cpu().numpy().astype(np.int64), minlength=209) / G
    p = np.zeros(209)
    p[la] = c[:len(la)] / c.sum()
    assert np.abs(freq - p).max() < 4 * np.sqrt(0.25 / G) and freq[p == 0].sum() == 0
    # other seeds / plies give other draws
    out_b = _advance(packed, rep(c), rep(a), torch.full((G,), len(la), dtype=torch.int16).cuda(), torch.arange(G).cuda(), 1.0, 100, 0, G)
    out_c = _advance(packed, rep(c), rep(a), torch.full((G,), len(la), dtype=torch.int16).cuda(), torch.arange(G).cuda(), 1.0, 99, 1, G)
    assert 0.5 < (out_b["chosen"] != out["chosen"]).float().mean() and 0.5 < (out_c["chosen"] != out["chosen"]).float().mean()


def test_flat_trainer_matches_torch_autograd_and_adam(traj):
    rng = np.random.default_rng(0)
    idx = rng.choice(len(traj["rows"]), 96, replace=False)
    rows = traj["rows"][idx]
    torch.manual_seed(3)
    pt = torch.softmax(2 * torch.randn(96, 209), 1).cuda()
    vt = torch.randint(-1, 2, (96,)).float().cuda()
    a = _net(1).train()
    b = copy.deepcopy(a)
    packed = gl.pack_rows(rows)
    trainer = train_network.FlatTrainer(a, lr=1e-3)
    opt = torch.optim.Adam(b.parameters(), lr=1e-3)
    ce, mse = torch.nn.CrossEntropyLoss(), torch.nn.MSELoss()
    for step in range(4):
        loss_flat = trainer.step(packed, pt, vt, 96).sum().item()
        p, v = b(packed)
        loss = ce(p, pt) + mse(v.squeeze(), vt)
        opt.zero_grad()
        loss.backward()
        opt.step()
        assert abs(loss_flat - loss.item()) <= 2e-5
    pa, pb = dict(a.named_parameters()), dict(b.named_parameters())
    # Adam divides by sqrt(v): where a gradient is ~0 (|g| << eps) its last-bit noise changes the update
    # by up to lr per step, so compare robustly: nearly all elements agree to 2e-6, none by more than 4*lr.
    # (aq_adam_step itself is checked to 2e-6 on well-scaled gradients in test_gpu_gnn.py.)
    for n in pa:
        d = (pa[n] - pb[n]).abs()
        assert d.max().item() <= 4e-3 and (d > 2e-6).float().mean().item() <= 0.02, (n, d.max().item())


def test_train_on_history_and_sharded_steps_equal_full_batch(traj):
    """Two half-batch steps whose gradients are summed (what the all-reduce does) give the same
    update as one full-batch step: data-parallel exactness on a single GPU."""
    rng = np.random.default_rng(1)
    idx = rng.choice(len(traj["rows"]), 64, replace=False)
    packed = gl.pack_rows(traj["rows"][idx])
    torch.manual_seed(4)
    pt = torch.softmax(torch.randn(64, 209), 1).cuda()
    vt = torch.randint(-1, 2, (64,)).float().cuda()
    full = train_network.FlatTrainer(_net(2).train())
    full.step(packed, pt, vt, 64)
    g_full = full.grads.clone()
    parts = []
    for rank in range(2):
        lo, hi = train_network.shard_bounds(64, rank, 2)
        t = train_network.FlatTrainer(_net(2).train())
        t.step(packed[lo:hi].contiguous(), pt[lo:hi].contiguous(), vt[lo:hi].contiguous(), 64)
        parts.append(t.grads.clone())
    assert (parts[0] + parts[1] - g_full).abs().max().item() <= 1e-6 * max(1.0, g_full.abs().max().item())
    # the full loop lowers the loss on a fixed history
    history = [[[r[0:2].tolist(), r[2:4].tolist(), r[4:].tolist()], pt[i].tolist(), float(vt[i])]
               for i, r in enumerate(traj["rows"][idx])]
    losses = train_network.train_on_history(_net(5), history, num_epochs=6, batch_size=32, verbose=False)
    assert sum(losses[-1]) < sum(losses[0])


def test_arena_and_tiny_train_cycle(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    model_dir = str(tmp_path / "models") + os.sep
    create_network(model_dir + "best.pth")
    assert os.path.exists(model_dir + "best.pth")
    monkeypatch.setattr(pv_mcts, "PV_EVALUATE_COUNT", 6)
    path = self_play.self_play(game_count=4, model_path=model_dir + "best.pth", data_dir=str(tmp_path / "data"), seed=3)
    hist = pickle.load(open(path, "rb"))
    assert len(hist) > 8 and len(hist[0]) == 3 and len(hist[0][1]) == 209
    train_network.train_network(data_dir=str(tmp_path / "data"), model_dir=model_dir, num_epochs=2)
    assert os.path.exists(model_dir + "latest.pth")
    promoted = evaluate_network.evaluate_network(model_dir=model_dir, num_games=4, sims=6)
    assert promoted in (True, False)
    pts = evaluate_network.play_matches(_net(0), _net(0), num_games=4, sims=6, seed=1)
    assert 0.0 <= pts <= 4.0


def _arena_evaluator(kind, salt):
    """The evaluators of tests/golden/make_arena_golden.py in torch integer ops (bit-identical priors and values)."""
    def ev(packed):
        rows, plies = gl.unpack_rows(packed)
        w = (torch.arange(68, device=rows.device, dtype=torch.int64) + 1) * 7919
        key = ((rows.to(torch.int64) * w).sum(1) + plies.to(torch.int64) * 104729) % (2 ** 31)
        mask, pawn = gl.legal_mask_batch(packed)
        dense = gl.mask_to_dense(mask)
        a = torch.arange(209, device=rows.device, dtype=torch.int64)
        if kind == "runner":
            prow, arow = (rows[:, 0].to(torch.int64) // 9)[:, None], (a // 9)[None, :]
            base = torch.where(arow < prow, 60, torch.where(arow == prow, 6, 2))
            base = torch.where(a[None, :] < 81, base, torch.ones_like(base))
            raw = base * 8 + (key[:, None] + a[None, :] * 40503 + salt) % 7
            val = torch.zeros(packed.shape[0], dtype=torch.float32, device=rows.device)
        else:
            raw = (key[:, None] + a[None, :] * 40503 + salt) % 1009 + 1
            val = (((key + salt) % 2001) - 1000).to(torch.float32) / torch.tensor(1000.0, dtype=torch.float32, device=rows.device)
        raw = torch.where(dense, raw, torch.zeros_like(raw))
        pri = raw.to(torch.float32) / raw.sum(1).to(torch.float32)[:, None]
        return {"priors": pri.contiguous(), "value": val.contiguous(), "mask": mask, "pawn": pawn}
    return ev


def test_arena_matches_reference_play_and_scoring():
    """evaluate_network.play / first_player_point / the colour alternation of the match loop (evaluate_network.py:18-45, 66-73):
    the UNMODIFIED reference played these pairings with deterministic evaluators at temperature 0
    (tests/golden/make_arena_golden.py); play_matches must reproduce every game move for move, every first-player point
    and the total."""
    import json
    with open(os.path.join(os.path.dirname(__file__), "golden", "arena_golden.json")) as f:
        golden = json.load(f)
    for pairing in golden["pairings"]:
        m0, m1 = _arena_evaluator(*pairing["model0"]), _arena_evaluator(*pairing["model1"])
        n = len(pairing["games"])
        total, points, log = evaluate_network.play_matches(m0, m1, num_games=n, temperature=golden["temperature"],
                                                           sims=golden["sims"], seed=0, details=True)
        for i, g in enumerate(pairing["games"]):
            assert log[i] == g["actions"], (pairing["name"], i)
            assert points[i] == g["first_player_point"], (pairing["name"], i)
        assert total == pairing["total_point_model0"], pairing["name"]
    # the fixture exercises what it should: both colours win somewhere, a draw occurs, and alternation changes the total
    pts = [g["first_player_point"] for p in golden["pairings"] for g in p["games"]]
    assert {0, 1, 0.5} <= set(pts)
    rr = [p for p in golden["pairings"] if p["name"] == "runner_vs_runner"][0]
    assert rr["total_point_model0"] == 2.0 and len(rr["games"]) == 5   # second mover always wins: model0 scores on odd games only


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_fused_training_step_equals_the_unfused_kernels(traj, prec):
    """FlatTrainer's default step (loss gradient inside the heads backward; slot reduction + all-reduce + Adam in one kernel with the
    bias corrections from the device step counter; the step replayed as a CUDA graph from the third call on) against the same step
    made of the separate entry points (aq_loss_grad, aq_gnn_backward, aq_adam_step): same losses and gradients bit for bit, same
    parameters up to the rounding of the bias-correction scalars."""
    rng = np.random.default_rng(2)
    idx = rng.choice(len(traj["rows"]), 200, replace=False)
    packed = gl.pack_rows(traj["rows"][idx])
    torch.manual_seed(4)
    pt = torch.softmax(2 * torch.randn(200, 209), 1).cuda()
    vt = torch.randint(-1, 2, (200,)).float().cuda()
    a, b, c = _net(7).train(), _net(7).train(), _net(7).train()
    fused = train_network.FlatTrainer(a, lr=1e-3, precision=prec)
    eager = train_network.FlatTrainer(c, lr=1e-3, precision=prec, use_graph=False)
    plain = train_network.FlatTrainer(b, lr=1e-3, precision=prec, collective="nccl")
    assert fused.collective == "p2p"
    for step in range(6):
        sl = slice(0, 200) if step != 3 else slice(0, 77)   # a second batch shape in between
        n = sl.stop
        lf = fused.step(packed[sl], pt[sl], vt[sl], n).clone()
        le = eager.step(packed[sl], pt[sl], vt[sl], n).clone()
        lp = plain.step(packed[sl], pt[sl], vt[sl], n).clone()
        # the loss scalars (policy loss ~ 5.3) are accumulated with float atomics, one per CTA: the summation order, hence the last
        # few ulps (4.8e-7 each at that magnitude), differs from launch to launch
        assert (lf - le).abs().max().item() <= 1e-5
        assert (lf - lp).abs().max().item() <= 1e-5
        assert torch.equal(fused.grads, eager.grads)
        assert (fused.grads - plain.grads).abs().max().item() <= 1e-6 * max(1.0, plain.grads.abs().max().item())
        assert (fused.flat - plain.flat).abs().max().item() <= 2e-6
        assert torch.equal(fused.flat, eager.flat)
    fused.check()
    assert fused.comm.status() == (6, 0)
    # an empty shard takes part in the step with a zero gradient
    fused.step(packed[:0], pt[:0], vt[:0], 64)
    plain.step(packed[:0], pt[:0], vt[:0], 64)
    assert (fused.flat - plain.flat).abs().max().item() <= 4e-6
