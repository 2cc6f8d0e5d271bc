"""CPU oracle for the integer half of the hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Two independent restatements of ``/root/reference/game_logic.py``:

* ``PyOracleState`` -- a small pure-Python restatement (lists and loops, like the
  reference) for small cases and as the "reference-speed" CPU baseline;
* the ctypes front-end of ``quoridor_oracle.c`` (``legal_actions_batch`` ...) -- the fast
  checker used for 10^4..10^6 positions.

Both are pinned against fixtures generated from the UNMODIFIED reference
(``tests/golden/make_golden.py``); see ``tests/test_oracle_golden.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package never does.

State interchange format ("row68"): ``uint8[68] = [player_pos, player_walls, enemy_pos,
enemy_walls, walls[64]]`` plus ``plies`` as a separate int16 (see quoridor_oracle.c).
"""
import ctypes
import os
import subprocess
from collections import deque

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libquoridor_oracle.so")
MAXACT = 136
DIRS = ((-1, 0), (1, 0), (0, -1), (0, 1))  # U, D, L, R -- game_logic.py:11


def build(force=False):
    """Compile quoridor_oracle.c with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "quoridor_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libquoridor_oracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, ll, i = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
        L.oq_legal_actions_batch.argtypes = [vp, vp, ll, i, vp, vp, vp, vp, i]
        L.oq_legal_actions_batch.restype = None
        L.oq_next_batch.argtypes = [vp, vp, vp, ll, i, i, vp, vp, vp]
        L.oq_next_batch.restype = None
        L.oq_open_mask_batch.argtypes = [vp, ll, i, vp]
        L.oq_open_mask_batch.restype = None
        L.oq_planes_batch.argtypes = [vp, ll, i, vp]
        L.oq_planes_batch.restype = None
        L.oq_legal_actions_pos_row.argtypes = [vp, i, i, vp]
        L.oq_legal_actions_pos_row.restype = i
        L.oq_legal_actions_wall_row.argtypes = [vp, i, i, vp]
        L.oq_legal_actions_wall_row.restype = i
        L.oq_max_threads.restype = i
        L.oq_heuristic_batch.argtypes = [vp, ll, i, i, i, vp, vp]
        L.oq_heuristic_batch.restype = None
        L.oq_alpha_beta_action.argtypes = [vp, i, i, i, i, i, vp]
        L.oq_alpha_beta_action.restype = i
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _rows(rows):
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    assert rows.ndim == 2 and rows.shape[1] == 68
    return rows


def max_threads():
    return lib().oq_max_threads()


def legal_actions_batch(rows, plies=None, N=9, nthreads=0):
    """-> dict(actions int16[M,136] (-1 padded, reference order), n int16[M],
    mask uint32[M,8], pawn uint8[M,8])."""
    rows = _rows(rows)
    M = rows.shape[0]
    pl = None if plies is None else np.ascontiguousarray(plies, dtype=np.int16)
    actions = np.empty((M, MAXACT), np.int16)
    n = np.empty(M, np.int16)
    mask = np.empty((M, 8), np.uint32)
    pawn = np.empty((M, 8), np.uint8)
    lib().oq_legal_actions_batch(_ptr(rows), _ptr(pl), M, N, _ptr(actions), _ptr(n), _ptr(mask),
                                 _ptr(pawn), nthreads)
    return {"actions": actions, "n": n, "mask": mask, "pawn": pawn}


def legal_mask_only(rows, N=9, nthreads=0):
    """Mask + ordered pawn list only (what the CUDA kernel emits); used for timing."""
    rows = _rows(rows)
    M = rows.shape[0]
    mask = np.empty((M, 8), np.uint32)
    pawn = np.empty((M, 8), np.uint8)
    lib().oq_legal_actions_batch(_ptr(rows), None, M, N, None, None, _ptr(mask), _ptr(pawn), nthreads)
    return mask, pawn


def next_batch(rows, plies, actions, N=9, plies_for_draw=116):
    rows = _rows(rows)
    M = rows.shape[0]
    pl = np.ascontiguousarray(plies, dtype=np.int16)
    ac = np.ascontiguousarray(actions, dtype=np.int16)
    out = np.zeros((M, 68), np.uint8)
    pl_out = np.empty(M, np.int16)
    flags = np.empty(M, np.uint8)
    lib().oq_next_batch(_ptr(rows), _ptr(pl), _ptr(ac), M, N, plies_for_draw, _ptr(out), _ptr(pl_out),
                        _ptr(flags))
    return out, pl_out, flags


def open_mask_batch(rows, N=9):
    rows = _rows(rows)
    out = np.empty((rows.shape[0], N * N), np.uint8)
    lib().oq_open_mask_batch(_ptr(rows), rows.shape[0], N, _ptr(out))
    return out


def planes_batch(rows, N=9):
    rows = _rows(rows)
    out = np.empty((rows.shape[0], 6, N, N), np.float32)
    lib().oq_planes_batch(_ptr(rows), rows.shape[0], N, _ptr(out))
    return out


def heuristic_batch(rows, N=9, plies_for_draw=116, num_walls=10):
    """agents.heuristic_eval (agents.py:22-54) -> (dist int16[M,2] = {mover, enemy} shortest paths, heur float64[M])."""
    rows = _rows(rows)
    dist = np.empty((rows.shape[0], 2), np.int16)
    heur = np.empty(rows.shape[0], np.float64)
    lib().oq_heuristic_batch(_ptr(rows), rows.shape[0], N, plies_for_draw, num_walls, _ptr(dist), _ptr(heur))
    return dist, heur


def alpha_beta_action(row, plies=0, max_depth=2, N=9, plies_for_draw=116, num_walls=10):
    """agents.alpha_beta_action (agents.py:90-107), literal fail-hard alpha-beta -> (action, root scores float64[n])."""
    row = np.ascontiguousarray(row, dtype=np.uint8)
    scores = np.full(MAXACT, np.nan, np.float64)
    a = lib().oq_alpha_beta_action(_ptr(row), int(plies), N, int(max_depth), plies_for_draw, num_walls, _ptr(scores))
    return a, scores


def legal_actions_pos(row, pos, N=9):
    row = np.ascontiguousarray(row, dtype=np.uint8)
    out = np.zeros(8, np.int32)
    k = lib().oq_legal_actions_pos_row(_ptr(row), N, int(pos), _ptr(out))
    return out[:k].tolist()


def legal_actions_wall(row, pos, N=9):
    row = np.ascontiguousarray(row, dtype=np.uint8)
    out = np.zeros(4, np.int32)
    k = lib().oq_legal_actions_wall_row(_ptr(row), N, int(pos), _ptr(out))
    return out[:k].tolist()


def edge_index_from_open(open_mask, N=9):
    """Directed edge list in canonical order: for node v ascending, neighbours in U,D,L,R
    order.  open_mask: uint8[V] of one board.  -> int64[2,E] (source row 0, target row 1)."""
    src, dst = [], []
    for v in range(N * N):
        for d, (dx, dy) in enumerate(DIRS):
            if (int(open_mask[v]) >> d) & 1:
                src.append(v)
                dst.append(v + dx * N + dy)
    return np.array([src, dst], dtype=np.int64)


def to_row68(player, enemy, walls):
    """State.to_array() triple -> row68."""
    row = np.zeros(68, np.uint8)
    row[0], row[1], row[2], row[3] = player[0], player[1], enemy[0], enemy[1]
    row[4:4 + len(walls)] = walls
    return row


# --------------------------------------------------------------------------------------
# Pure-Python restatement (small cases / reference-speed baseline)
# --------------------------------------------------------------------------------------
class PyOracleState:
    """Restates game_logic.State (game_logic.py:15-395) with the same constructor
    arguments and attribute names, so it can stand in for the reference State in the
    reference's own pv_mcts loop on boxes where /root/reference is absent."""

    PLIES_FOR_DRAW = {3: 14, 5: 28, 9: 116}  # constants.py:6-20

    def __init__(self, board_size=9, num_walls=10, player=None, enemy=None, walls=None, plies_played=0):
        if board_size % 2 == 0:
            raise ValueError("The board size must be an odd number.")
        self.N = board_size
        start = board_size * (board_size - 1) + board_size // 2  # game_logic.py:36
        fresh = player is None or enemy is None
        self.player = [start, num_walls] if fresh else player
        self.enemy = [start, num_walls] if fresh else enemy
        self.walls = walls if walls is not None else [0] * ((board_size - 1) ** 2)
        self.plies_played = plies_played

    # game_logic.py:43-54
    def is_lose(self):
        return self.enemy[0] // self.N == 0

    def is_draw(self):
        return self.plies_played >= self.PLIES_FOR_DRAW[self.N]

    def is_done(self):
        return self.is_lose() or self.is_draw()

    def is_first_player(self):  # game_logic.py:394-395
        return self.plies_played % 2 == 0

    def to_array(self):  # game_logic.py:96-100
        return [list(self.player), list(self.enemy), list(self.walls)]

    def _blocked(self, x, y, nx, ny):  # game_logic.py:145-167
        N, M, w = self.N, self.N - 1, self.walls
        if nx != x:
            r = min(x, nx)  # wall row between the two squares
            return (y < N - 1 and w[r * M + y] == 1) or (y > 0 and w[r * M + y - 1] == 1)
        if ny != y:
            c = min(y, ny)  # wall column between the two squares
            return (x < N - 1 and w[x * M + c] == 2) or (x > 0 and w[(x - 1) * M + c] == 2)
        return False

    def legal_actions_pos(self, pos):  # game_logic.py:120-192
        N = self.N
        inside = lambda a, b: 0 <= a < N and 0 <= b < N
        x, y = divmod(pos, N)
        ex, ey = divmod(N * N - 1 - self.enemy[0], N)
        moves = []
        for dx, dy in DIRS:
            nx, ny = x + dx, y + dy
            if not inside(nx, ny) or self._blocked(x, y, nx, ny):
                continue
            if (nx, ny) != (ex, ey):
                moves.append(nx * N + ny)
                continue
            jx, jy = nx + dx, ny + dy
            if inside(jx, jy) and not self._blocked(nx, ny, jx, jy):
                moves.append(jx * N + jy)
                continue
            sides = ((0, -1), (0, 1)) if dx != 0 else ((-1, 0), (1, 0))
            for sx, sy in sides:
                tx, ty = nx + sx, ny + sy
                if inside(tx, ty) and not self._blocked(nx, ny, tx, ty):
                    moves.append(tx * N + ty)
        return moves

    def _can_place(self, o, pos):  # game_logic.py:199-223
        N, M, w = self.N, self.N - 1, self.walls
        if w[pos] != 0:
            return False
        x, y = divmod(pos, M)
        if o == 1:
            return not ((y > 0 and w[pos - 1] == 1) or (y < N - 2 and w[pos + 1] == 1))
        return not ((x > 0 and w[pos - M] == 2) or (x < N - 2 and w[pos + M] == 2))

    def _gate(self, o, pos):  # game_logic.py:227-307
        N, M, w = self.N, self.N - 1, self.walls
        x, y = divmod(pos, M)
        if o == 1:
            along, across, step, cross = y, x, 1, M
        else:
            along, across, step, cross = x, y, M, 1
        other = 3 - o  # perpendicular orientation touches the ends/middle

        def perp_at(p):  # any perpendicular wall at p or its two neighbours across
            return (w[p] == other or (across > 0 and w[p - cross] == other)
                    or (across < N - 2 and w[p + cross] == other))

        first = along == 0 or perp_at(pos - step) or (along > 1 and w[pos - 2 * step] == o)
        middle = (across > 0 and w[pos - cross] == other) or (across < N - 2 and w[pos + cross] == other)
        last = along == N - 2 or perp_at(pos + step) or (along < N - 3 and w[pos + 2 * step] == o)
        return first + middle + last >= 2

    def _bfs(self):  # game_logic.py:309-324
        seen = {self.player[0]}
        todo = deque(seen)
        while todo:
            p = todo.popleft()
            if p // self.N == 0:
                return True
            for q in self.legal_actions_pos(p):
                if q not in seen:
                    seen.add(q)
                    todo.append(q)
        return False

    def legal_actions_wall(self, pos):  # game_logic.py:195-357
        N = self.N
        out = []
        for o, base in ((1, N * N), (2, N * N + (N - 1) ** 2)):
            if not self._can_place(o, pos):
                continue
            ok = True
            if self._gate(o, pos):
                trial = PyOracleState(N, 0, list(self.player), list(self.enemy), list(self.walls), self.plies_played)
                trial.walls[pos] = o
                ok = trial._bfs() and trial.next(base + pos)._bfs()
            if ok:
                out.append(base + pos)
        return out

    def legal_actions(self):  # game_logic.py:103-117
        acts = self.legal_actions_pos(self.player[0])
        if self.player[1] > 0:
            for pos in range((self.N - 1) ** 2):
                acts.extend(self.legal_actions_wall(pos))
        return acts

    def next(self, action):  # game_logic.py:359-391
        N = self.N
        S = (N - 1) ** 2
        me, walls = list(self.player), list(self.walls)
        if action < N * N:
            me[0] = action
        else:
            slot, o = (action - N * N, 1) if action < N * N + S else (action - N * N - S, 2)
            walls[slot] = o
            me[1] -= 1
        return PyOracleState(N, 0, list(self.enemy), me, walls[::-1], self.plies_played + 1)

    def row68(self):
        return to_row68(self.player, self.enemy, self.walls)

    @classmethod
    def from_row68(cls, row, plies=0, N=9):
        row = [int(v) for v in row]
        return cls(N, 0, row[0:2], row[2:4], row[4:4 + (N - 1) ** 2], int(plies))
