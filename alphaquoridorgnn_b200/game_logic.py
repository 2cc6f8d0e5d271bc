"""Quoridor game logic backed by the CUDA kernels (mirror of reference game_logic.py).

Two levels:

* batched device API (what the hot path uses): ``pack_rows``, ``legal_mask_batch``,
  ``legal_actions_batch``, ``next_batch``, ``build_graph_batch`` operate on packed states, a
  ``uint8[B,32]`` CUDA tensor holding one ``AqState`` (include/aqgnn.h) per row;
* ``State``: the reference's class (game_logic.py:15-395) with the same constructor, attributes
  and methods, so pv_mcts.py / self_play.py style callers keep working.  Record keeping (next,
  is_lose, to_array ...) is plain host code; everything that searches (legal_actions,
  legal_actions_pos, legal_actions_wall) runs the CUDA kernel.  There is no CPU fallback.
"""
import numpy as np
import torch

from . import _lib
from .constants import BOARD_SIZE, NUM_WALLS, NUM_PLIES_FOR_DRAW

N = 9
NUM_SQUARES = 81
NUM_SLOTS = 64
NUM_ACTIONS = NUM_SQUARES + 2 * NUM_SLOTS  # 209
MAX_LEGAL = 136
STATE_BYTES = 32


def _dev(device=None):
    if not torch.cuda.is_available():
        raise _lib.AqError("no CUDA device: the AlphaQuoridorGNN hot path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


# ---------------------------------------------------------------------------------------------
# host-side conversion of python State objects / to_array() triples to row68
# ---------------------------------------------------------------------------------------------
def rows_from_arrays(game_state_arrays):
    """list of [player, enemy, walls] (State.to_array(), game_logic.py:96-100) -> uint8[B,68]."""
    B = len(game_state_arrays)
    rows = np.zeros((B, 68), np.uint8)
    for i, (player, enemy, walls) in enumerate(game_state_arrays):
        rows[i, 0], rows[i, 1], rows[i, 2], rows[i, 3] = player[0], player[1], enemy[0], enemy[1]
        if len(walls) != NUM_SLOTS:
            raise ValueError("the CUDA kernels are built for the 9x9 board (64 wall slots)")
        rows[i, 4:] = walls
    return rows


def rows_from_states(states):
    rows = rows_from_arrays([s.to_array() for s in states])
    plies = np.array([s.plies_played for s in states], np.int16)
    return rows, plies


# ---------------------------------------------------------------------------------------------
# batched device API
# ---------------------------------------------------------------------------------------------
def pack_rows(rows, plies=None, device=None):
    """uint8[B,68] (+ int16[B] plies) -> packed uint8[B,32] CUDA tensor (aq_pack_states)."""
    dev = _dev(device)
    rows = torch.as_tensor(rows, dtype=torch.uint8).to(dev).contiguous()
    B = rows.shape[0]
    pl = None if plies is None else torch.as_tensor(plies, dtype=torch.int16).to(dev).contiguous()
    out = torch.empty((B, STATE_BYTES), dtype=torch.uint8, device=dev)
    L = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(L.aq_pack_states(_lib.ptr(rows), _lib.ptr(pl), B, _lib.ptr(out), _lib.stream_ptr(dev)), "aq_pack_states")
    return out


def unpack_rows(packed):
    _lib.require_cuda(packed, "packed")
    B = packed.shape[0]
    rows = torch.empty((B, 68), dtype=torch.uint8, device=packed.device)
    plies = torch.empty((B,), dtype=torch.int16, device=packed.device)
    L = _lib.load()
    with torch.cuda.device(packed.device):
        _lib.check(L.aq_unpack_states(_lib.ptr(packed), B, _lib.ptr(rows), _lib.ptr(plies), _lib.stream_ptr(packed.device)),
                   "aq_unpack_states")
    return rows, plies


def pack_rows_host(rows, plies=None):
    """Host-side packing (numpy) into the AqState byte layout; used to fill pinned host buffers for
    the host-buffer entry points.  Pure data-layout conversion, no game logic."""
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    B = rows.shape[0]
    walls = rows[:, 4:]
    weights = (np.uint64(1) << np.arange(64, dtype=np.uint64))
    out = np.zeros((B, 4), np.uint64)
    out[:, 0] = ((walls == 1).astype(np.uint64) * weights).sum(axis=1, dtype=np.uint64)
    out[:, 1] = ((walls == 2).astype(np.uint64) * weights).sum(axis=1, dtype=np.uint64)
    meta = (rows[:, 0].astype(np.uint64) | (rows[:, 1].astype(np.uint64) << np.uint64(8))
            | (rows[:, 2].astype(np.uint64) << np.uint64(16)) | (rows[:, 3].astype(np.uint64) << np.uint64(24)))
    if plies is not None:
        meta |= (np.asarray(plies).astype(np.uint64) & np.uint64(0xFFFF)) << np.uint64(32)
    out[:, 2] = meta
    return out.view(np.uint8).reshape(B, STATE_BYTES)


def legal_mask_batch(packed):
    """State.legal_actions() for every row: -> (mask int32[B,8] bitmask over the 209 actions,
    pawn uint8[B,8] = [n, ordered pawn moves..., pad])."""
    _lib.require_cuda(packed, "packed")
    B = packed.shape[0]
    mask = torch.empty((B, 8), dtype=torch.int32, device=packed.device)
    pawn = torch.empty((B, 8), dtype=torch.uint8, device=packed.device)
    L = _lib.load()
    ws = torch.empty((L.aq_legal_mask_ws_bytes(B),), dtype=torch.uint8, device=packed.device)  # task list of the two-phase form
    with torch.cuda.device(packed.device):
        _lib.check(L.aq_legal_mask_ws(_lib.ptr(packed), B, _lib.ptr(mask), _lib.ptr(pawn), _lib.ptr(ws), ws.numel(),
                                      _lib.stream_ptr(packed.device)), "aq_legal_mask_ws")
    return mask, pawn


def legal_actions_batch(packed, mask=None, pawn=None):
    """-> (actions int16[B,136] in State.legal_actions() order, -1 padded; n int16[B])."""
    if mask is None:
        mask, pawn = legal_mask_batch(packed)
    B = mask.shape[0]
    actions = torch.empty((B, MAX_LEGAL), dtype=torch.int16, device=mask.device)
    n = torch.empty((B,), dtype=torch.int16, device=mask.device)
    L = _lib.load()
    with torch.cuda.device(mask.device):
        _lib.check(L.aq_legal_actions_list(_lib.ptr(mask), _lib.ptr(pawn), B, _lib.ptr(actions), _lib.ptr(n),
                                           _lib.stream_ptr(mask.device)), "aq_legal_actions_list")
    return actions, n


def mask_to_dense(mask):
    """int32[B,8] bitmask -> bool[B,209]."""
    bits = torch.arange(32, device=mask.device, dtype=torch.int32)
    dense = ((mask.unsqueeze(-1) >> bits) & 1).reshape(mask.shape[0], 256)
    return dense[:, :NUM_ACTIONS].bool()


def next_batch(packed, actions):
    """State.next(action) per row -> (packed', terminal uint8[B]: bit0 is_lose, bit1 is_draw)."""
    _lib.require_cuda(packed, "packed")
    B = packed.shape[0]
    actions = actions.to(device=packed.device, dtype=torch.int16).contiguous()
    out = torch.empty_like(packed)
    term = torch.empty((B,), dtype=torch.uint8, device=packed.device)
    L = _lib.load()
    with torch.cuda.device(packed.device):
        _lib.check(L.aq_state_next(_lib.ptr(packed), _lib.ptr(actions), B, _lib.ptr(out), _lib.ptr(term),
                                   _lib.stream_ptr(packed.device)), "aq_state_next")
    return out, term


def build_graph_batch(packed, with_edge_index=False):
    """Board graph of every state: dict(open_mask uint8[B,81], dinv f32[B,81], x f32[B*81,6],
    and with_edge_index: edge_index int64[2,E], batch int64[B*81])."""
    _lib.require_cuda(packed, "packed")
    dev = packed.device
    B = packed.shape[0]
    nbytes = (B * NUM_SQUARES + 3) // 4 * 4
    open_mask = torch.empty((nbytes,), dtype=torch.uint8, device=dev)[: B * NUM_SQUARES].view(B, NUM_SQUARES)
    dinv = torch.empty((B, NUM_SQUARES), dtype=torch.float32, device=dev)
    x = torch.empty((B * NUM_SQUARES, 6), dtype=torch.float32, device=dev)
    cnt = torch.empty((B,), dtype=torch.int32, device=dev)
    L = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(L.aq_build_graph(_lib.ptr(packed), B, _lib.ptr(open_mask), _lib.ptr(dinv), _lib.ptr(x), _lib.ptr(cnt),
                                    _lib.stream_ptr(dev)), "aq_build_graph")
        out = {"open_mask": open_mask, "dinv": dinv, "x": x, "edge_count": cnt}
        if with_edge_index:
            cs = torch.cumsum(cnt.to(torch.int64), 0)
            off = (cs - cnt).contiguous()
            E = int(cs[-1].item()) if B > 0 else 0
            ei = torch.empty((2, E), dtype=torch.int64, device=dev)
            _lib.check(L.aq_build_edge_index(_lib.ptr(open_mask), _lib.ptr(off), B, _lib.ptr(ei[0]), _lib.ptr(ei[1]),
                                             _lib.stream_ptr(dev)), "aq_build_edge_index")
            out["edge_index"] = ei
            out["batch"] = torch.arange(B, device=dev, dtype=torch.int64).repeat_interleave(NUM_SQUARES)
    return out


def open_mask_from_edge_index(edge_index, num_graphs):
    """(x, edge_index, batch) callers: recover per-node open-direction masks; raises if the edges
    are not 4-neighbour edges of 9x9 boards."""
    _lib.require_cuda(edge_index, "edge_index")
    dev = edge_index.device
    ei = edge_index.to(torch.int64).contiguous()
    nbytes = (num_graphs * NUM_SQUARES + 3) // 4 * 4
    buf = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    bad = torch.zeros((1,), dtype=torch.int32, device=dev)
    L = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(L.aq_edges_to_open_mask(_lib.ptr(ei[0]), _lib.ptr(ei[1]), ei.shape[1], num_graphs, _lib.ptr(buf),
                                           _lib.ptr(bad), _lib.stream_ptr(dev)), "aq_edges_to_open_mask")
    code = int(bad.item())
    if code == 2:
        raise ValueError("edge_index is not symmetric: every edge of a board graph has its reverse (GCNConv's target-degree "
                         "normalisation and the kernels' node-degree normalisation agree only then)")
    if code != 0:
        raise ValueError("edge_index is not a 9x9 Quoridor board graph (edges must join 4-neighbours of one board)")
    return buf[: num_graphs * NUM_SQUARES].view(num_graphs, NUM_SQUARES)


# ---------------------------------------------------------------------------------------------
# reference-compatible State
# ---------------------------------------------------------------------------------------------
class State:
    """Same constructor and attributes as the reference State (game_logic.py:25-40)."""

    def __init__(self, board_size=BOARD_SIZE, num_walls=NUM_WALLS, player=None, enemy=None, walls=None, plies_played=0):
        if board_size % 2 == 0:
            raise ValueError('The board size must be an odd number.')
        if board_size != N:
            raise ValueError('the CUDA kernels are built for board_size=9')
        self.N = board_size
        self.player = player if player is not None else [0] * 2
        self.enemy = enemy if enemy is not None else [0] * 2
        self.walls = walls if walls is not None else [0] * NUM_SLOTS
        self.plies_played = plies_played
        if player is None or enemy is None:
            init_pos = N * (N - 1) + N // 2
            self.player[0] = init_pos
            self.player[1] = num_walls
            self.enemy[0] = init_pos
            self.enemy[1] = num_walls

    # ---- record keeping (host) ---------------------------------------------------------------
    def is_lose(self):  # game_logic.py:43-46
        return self.enemy[0] // self.N == 0

    def is_draw(self):  # game_logic.py:49-50
        return self.plies_played >= NUM_PLIES_FOR_DRAW

    def is_done(self):
        return self.is_lose() or self.is_draw()

    def is_first_player(self):  # game_logic.py:394-395
        return self.plies_played % 2 == 0

    def to_array(self):  # game_logic.py:96-100
        return [list(self.player), list(self.enemy), list(self.walls)]

    def pieces_array(self):  # game_logic.py:56-93
        g = build_graph_batch(self._packed())
        planes = g["x"].view(NUM_SQUARES, 6).t().cpu().numpy().astype(np.int64)
        return [[planes[0].tolist(), planes[1].tolist()], [planes[2].tolist(), planes[3].tolist()],
                [planes[4].tolist(), planes[5].tolist()]]

    def rotate_walls(self):  # game_logic.py:359-364
        self.walls = list(self.walls[::-1])

    def next(self, action):  # game_logic.py:366-391
        mover, walls = list(self.player), list(self.walls)
        if action < NUM_SQUARES:
            mover[0] = action
        elif action < NUM_SQUARES + NUM_SLOTS:
            walls[action - NUM_SQUARES] = 1
            mover[1] -= 1
        else:
            walls[action - NUM_SQUARES - NUM_SLOTS] = 2
            mover[1] -= 1
        return State(self.N, player=list(self.enemy), enemy=mover, walls=walls[::-1], plies_played=self.plies_played + 1)

    # ---- searches (CUDA) -----------------------------------------------------------------------
    def _packed(self, player=None):
        rows = rows_from_arrays([[player if player is not None else self.player, self.enemy, self.walls]])
        return pack_rows(rows, np.array([self.plies_played], np.int16))

    def legal_actions(self):  # game_logic.py:103-117
        actions, n = legal_actions_batch(self._packed())
        return actions[0, : int(n[0])].tolist()

    def legal_actions_pos(self, pos):  # game_logic.py:120-192
        _, pawn = legal_mask_batch(self._packed(player=[pos, 0]))
        p = pawn[0].tolist()
        return p[1:1 + p[0]]

    def legal_actions_wall(self, pos):  # game_logic.py:195-357 (ignores walls in hand, like the reference)
        mask, _ = legal_mask_batch(self._packed(player=[self.player[0], max(1, self.player[1])]))
        dense = mask_to_dense(mask)[0]
        return [a for a in (NUM_SQUARES + pos, NUM_SQUARES + NUM_SLOTS + pos) if bool(dense[a])]
