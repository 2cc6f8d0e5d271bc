"""GPU parity tests of the integer path (legal mask, ordered actions, next, graph) through the
C ABI: bit-exact against the golden fixtures (reference output) and against the CPU oracle."""
import numpy as np
import pytest
import torch

from alphaquoridorgnn_b200 import game_logic as gl
from alphaquoridorgnn_b200 import positions
from oracle import quoridor_oracle as qo

pytestmark = pytest.mark.gpu


def _np(t):
    return t.cpu().numpy()


def test_pack_unpack_roundtrip(traj):
    rows, plies = traj["rows"], traj["plies"]
    packed = gl.pack_rows(rows, plies)
    assert np.array_equal(_np(packed), gl.pack_rows_host(rows, plies))
    r2, p2 = gl.unpack_rows(packed)
    assert np.array_equal(_np(r2), rows) and np.array_equal(_np(p2), plies)


def test_legal_mask_golden_exact(traj):
    packed = gl.pack_rows(traj["rows"], traj["plies"])
    mask, pawn = gl.legal_mask_batch(packed)
    assert np.array_equal(_np(mask).view(np.uint32), traj["mask"])
    assert np.array_equal(_np(pawn), traj["pawn"])
    actions, n = gl.legal_actions_batch(packed, mask, pawn)
    assert np.array_equal(_np(n), traj["nact"])
    ref = qo.legal_actions_batch(traj["rows"], traj["plies"])
    assert np.array_equal(_np(actions), ref["actions"])


def test_legal_mask_ragged_and_empty():
    for B in (0, 1, 3, 5, 33):
        rows = np.zeros((B, 68), np.uint8)
        rows[:, 0] = 76; rows[:, 1] = 10; rows[:, 2] = 76; rows[:, 3] = 10
        packed = gl.pack_rows(rows, np.zeros(B, np.int16))
        mask, pawn = gl.legal_mask_batch(packed)
        actions, n = gl.legal_actions_batch(packed, mask, pawn)
        assert mask.shape == (B, 8) and actions.shape == (B, 136)
        if B:
            assert (_np(n) == 131).all()  # KA1
            assert _np(actions)[0, :7].tolist() == [67, 75, 77, 81, 145, 82, 146]


def test_state_next_golden(traj):
    rows, plies, game, action = traj["rows"], traj["plies"], traj["game"], traj["action"]
    idx = np.nonzero((game[:-1] >= 0) & (game[:-1] == game[1:]) & (action[:-1] >= 0))[0]
    packed = gl.pack_rows(rows[idx], plies[idx])
    nxt, term = gl.next_batch(packed, torch.from_numpy(action[idx]))
    r2, p2 = gl.unpack_rows(nxt)
    assert np.array_equal(_np(r2), rows[idx + 1]) and np.array_equal(_np(p2), plies[idx + 1])
    lose = (rows[idx + 1][:, 2] // 9) == 0
    draw = plies[idx + 1] >= 116
    assert np.array_equal(_np(term), lose.astype(np.uint8) | (draw.astype(np.uint8) << 1))


def test_graph_golden(traj, graph_golden, ka):
    idx = graph_golden["index"]
    packed = gl.pack_rows(traj["rows"][idx], traj["plies"][idx])
    g = gl.build_graph_batch(packed, with_edge_index=True)
    B = len(idx)
    assert np.array_equal(_np(g["open_mask"]), graph_golden["open"])
    planes = _np(g["x"]).reshape(B, 81, 6).transpose(0, 2, 1).reshape(B, 6, 9, 9)
    assert np.array_equal(planes, graph_golden["planes"])
    deg = 1 + np.unpackbits(graph_golden["open"][..., None], axis=-1).sum(-1)
    assert np.allclose(_np(g["dinv"]), deg.astype(np.float64) ** -0.5, rtol=0, atol=1e-7)
    # canonical edge list, symmetric, E = 288 - 4 * walls for legal positions
    ei = _np(g["edge_index"])
    want = np.concatenate([qo.edge_index_from_open(graph_golden["open"][b]) + 81 * b for b in range(B)], axis=1)
    assert np.array_equal(ei, want)
    fwd = set(map(tuple, ei.T.tolist()))
    assert all((t, s) in fwd for (s, t) in fwd)
    # round trip through the (x, edge_index, batch) entry
    om = gl.open_mask_from_edge_index(g["edge_index"], B)
    assert np.array_equal(_np(om), graph_golden["open"])
    # KA11 planes through the State class
    v = ka["KA11"]
    s = gl.State(player=v["row"][0:2], enemy=v["row"][2:4], walls=v["row"][4:])
    assert np.array_equal(np.array(s.pieces_array()).reshape(6, 9, 9), np.array(v["pieces_array"]))


def test_bad_edge_index_is_rejected():
    ei = torch.tensor([[0, 5], [1, 40]], dtype=torch.int64, device="cuda")
    with pytest.raises(ValueError):
        gl.open_mask_from_edge_index(ei, 1)


def test_ka_vectors_through_state_class(ka):
    g1 = ka["G1"]
    s = gl.State(player=g1["row"][0:2], enemy=g1["row"][2:4], walls=g1["row"][4:])
    assert s.legal_actions_wall(26) == []                                                    # G1
    for name in ("KA1", "KA2", "KA3", "KA4", "KA5", "KA6", "KA13", "G1"):
        v = ka[name]
        s = gl.State(player=v["row"][0:2], enemy=v["row"][2:4], walls=v["row"][4:], plies_played=v["plies"])
        assert s.legal_actions() == v["legal_actions"], name
        if "legal_actions_pos" in v:
            assert s.legal_actions_pos(v.get("pos", v["row"][0])) == v["legal_actions_pos"], name
    v = ka["KA10"]
    s = gl.State(player=v["row"][0:2], enemy=v["row"][2:4], walls=v["row"][4:])
    for p, want in v["legal_actions_wall"].items():
        assert s.legal_actions_wall(int(p)) == want


def test_gpu_generated_positions_match_oracle_at_scale():
    """Size-independent check at BASELINE config 2 scale: positions generated on the GPU by random
    legal continuation; the full set is compared with the C oracle (exact)."""
    packed = positions.random_positions(1_000_000, seed=0, games=16384)
    assert packed.shape == (1_000_000, 32)
    mask, pawn = gl.legal_mask_batch(packed)
    rows, plies = gl.unpack_rows(packed)
    rows, plies = _np(rows), _np(plies)
    ref_mask, ref_pawn = qo.legal_mask_only(rows)
    assert np.array_equal(_np(mask).view(np.uint32), ref_mask)
    assert np.array_equal(_np(pawn), ref_pawn)
    # walls in hand and walls on board are both well covered
    assert (rows[:, 1] > 0).mean() > 0.2 and ((rows[:, 4:] != 0).sum(1) >= 6).mean() > 0.3
    # idempotence / determinism
    mask2, pawn2 = gl.legal_mask_batch(packed)
    assert torch.equal(mask, mask2) and torch.equal(pawn, pawn2)
    # every legal action leads to a state whose wall count / pawn square is consistent
    actions, n = gl.legal_actions_batch(packed, mask, pawn)
    first = actions[:, 0].clone()
    nxt, term = gl.next_batch(packed, first)
    r2, p2 = gl.unpack_rows(nxt)
    ref_next, ref_plies, ref_flags = qo.next_batch(rows, plies, _np(first))
    assert np.array_equal(_np(r2), ref_next) and np.array_equal(_np(p2), ref_plies)
    assert np.array_equal(_np(term), ref_flags)


def test_legal_mask_every_form_agrees_with_reference(traj):
    """The one-kernel form at 2 / 8 / 32 lanes per state (chosen by the batch size: <= 1,024 states 32 lanes, <= 16,384 eight,
    above two), the two-phase form with its recommended workspace and with a workspace so small that most states overflow the
    task list and search in the first kernel: all bit-exact."""
    from alphaquoridorgnn_b200 import _lib
    L = _lib.load()
    packed = gl.pack_rows(traj["rows"], traj["plies"])
    B = packed.shape[0]
    want_mask, want_pawn = traj["mask"], traj["pawn"]

    def run(ws_bytes, lo=0, hi=B):
        n = hi - lo
        mask = torch.zeros((n, 8), dtype=torch.int32, device="cuda")
        pawn = torch.zeros((n, 8), dtype=torch.uint8, device="cuda")
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device="cuda") if ws_bytes else None
        _lib.check(L.aq_legal_mask_ws(_lib.ptr(packed[lo:hi]), n, _lib.ptr(mask), _lib.ptr(pawn), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()),
                   "aq_legal_mask_ws")
        assert np.array_equal(_np(mask).view(np.uint32), want_mask[lo:hi]) and np.array_equal(_np(pawn), want_pawn[lo:hi])

    run(0)                                  # one kernel, 2 lanes per state (54 k states)
    for lo in range(0, B, 9000):
        run(0, lo, min(B, lo + 9000))       # one kernel, 8 lanes per state
    for lo in range(0, B, 7001):
        run(0, lo, min(B, lo + 1000))       # one kernel, 32 lanes per state
    run(L.aq_legal_mask_ws_bytes(3000), 100, 3100)   # small batches with a workspace still take the one-kernel form
    run(L.aq_legal_mask_ws_bytes(B))        # two-phase, everything through the list
    run(256 + 4096)                         # 1,024 list entries for 54k states: overflow path
    run(256 + 4 * 30000)
    mask = torch.zeros((B, 8), dtype=torch.int32, device="cuda")
    pawn = torch.zeros((B, 8), dtype=torch.uint8, device="cuda")
    _lib.check(L.aq_legal_mask(_lib.ptr(packed), B, _lib.ptr(mask), _lib.ptr(pawn), _lib.stream_ptr()), "aq_legal_mask")
    assert np.array_equal(_np(mask).view(np.uint32), want_mask) and np.array_equal(_np(pawn), want_pawn)


def test_two_phase_legal_mask_back_to_back_calls_with_different_task_counts(traj):
    """Regression: the search kernel is launched programmatically behind the prepare kernel; its loads of the task count and the
    task list must come after its grid-dependency wait (a `const __restrict__` load was hoisted above it and read the PREVIOUS
    call's count).  Alternating batches with many / no / few path searches through the same workspace must all be exact."""
    from alphaquoridorgnn_b200 import _lib
    L = _lib.load()
    rows, plies = traj["rows"], traj["plies"]
    nwalls = (rows[:, 4:] != 0).sum(1)
    hard = np.argsort(-nwalls)[:6000]          # many walls on the board: many gated candidates
    easy = np.argsort(nwalls, kind="stable")[:6000]   # fewest walls: (almost) no path searches
    mid = np.arange(5000, 11000)
    ws = torch.empty((L.aq_legal_mask_ws_bytes(6000),), dtype=torch.uint8, device="cuda")
    packed = {k: gl.pack_rows(rows[idx], plies[idx]) for k, idx in (("hard", hard), ("easy", easy), ("mid", mid))}
    want = {k: (traj["mask"][idx], traj["pawn"][idx]) for k, idx in (("hard", hard), ("easy", easy), ("mid", mid))}
    outs = []
    for k in ("hard", "easy", "hard", "mid", "easy", "mid", "hard"):
        mask = torch.zeros((6000, 8), dtype=torch.int32, device="cuda")
        pawn = torch.zeros((6000, 8), dtype=torch.uint8, device="cuda")
        _lib.check(L.aq_legal_mask_ws(_lib.ptr(packed[k]), 6000, _lib.ptr(mask), _lib.ptr(pawn), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
                   "aq_legal_mask_ws")
        outs.append((k, mask, pawn))  # no synchronisation between the calls
    torch.cuda.synchronize()
    for k, mask, pawn in outs:
        assert np.array_equal(_np(mask).view(np.uint32), want[k][0]), k
        assert np.array_equal(_np(pawn), want[k][1]), k
