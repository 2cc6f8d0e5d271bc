"""Baseline agents backed by the CUDA kernels (mirror of reference agents.py).

* ``heuristic_eval(state)`` / ``heuristic_eval_batch(packed)`` -- agents.py:22-54: shortest-path difference,
  both searches in ``shortest_paths_kernel`` (one thread per state, 81-bit flood fills).
* ``alpha_beta_action(state, max_depth=2)`` -- agents.py:90-107.  The reference walks the tree one node at a time
  with fail-hard alpha-beta pruning; here the whole depth-limited tree is expanded level by level on the GPU
  (legal mask -> ordered action list -> ``state_next`` for every child at once), the depth-0 leaves are scored by
  ``aq_shortest_paths`` and every level is backed up by ``aq_negamax_backup``.  Pruning never changes the value of
  a node whose true value lies inside its (alpha, beta) window, and at the root a move is taken only when its
  score is STRICTLY greater than the best so far (agents.py:102-104), i.e. only when it was searched with its
  true value inside the window -- so the action chosen is the first arg-max of the exact negamax values, which is
  what the level-synchronous version computes.  Values are kept as exact integers (48 x the reference's floats:
  every leaf is -1, 0 or k/48).
* ``random_action(state)`` -- agents.py:14-18.

``mcts_action`` (agents.py:110-211: random-playout MCTS) is not mirrored: it is a sequential host loop of single
random playouts with no batch dimension and no deterministic output to compare.  There is no CPU fallback.
"""
import random

import torch

from . import _lib
from . import game_logic as gl
from .constants import NUM_PLIES_FOR_DRAW, NUM_WALLS

MAX_DIST_FROM_GOAL = NUM_PLIES_FOR_DRAW // 2 - NUM_WALLS  # agents.py:11
_NOT_FIXED = -(2 ** 31)


def random_action(state):
    """agents.py:14-18."""
    legal_actions = state.legal_actions()
    return legal_actions[random.randint(0, len(legal_actions) - 1)]


def shortest_paths_batch(packed, want_heuristic=True, want_leaf=False):
    """packed uint8[B,32] CUDA -> dict(dist int16[B,2] = {mover, enemy}, heuristic float64[B], leaf48 int32[B])."""
    _lib.require_cuda(packed, "packed")
    dev, B = packed.device, packed.shape[0]
    dist = torch.empty((B, 2), dtype=torch.int16, device=dev)
    heur = torch.empty((B,), dtype=torch.float64, device=dev) if want_heuristic else None
    leaf = torch.empty((B,), dtype=torch.int32, device=dev) if want_leaf else None
    L = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(L.aq_shortest_paths(_lib.ptr(packed), B, _lib.ptr(dist), _lib.ptr(heur), _lib.ptr(leaf), _lib.stream_ptr(dev)),
                   "aq_shortest_paths")
    return {"dist": dist, "heuristic": heur, "leaf48": leaf}


def heuristic_eval_batch(packed):
    """agents.heuristic_eval for every row -> float64[B] CUDA tensor."""
    return shortest_paths_batch(packed)["heuristic"]


def heuristic_eval(state):
    """agents.py:22-54 for one State (does not touch the state; the reference rotates it and rotates it back)."""
    return float(heuristic_eval_batch(state._packed())[0].item())


def _expand(packed):
    """Children of every non-terminal row in legal_actions() order.
    -> (children packed [C,32], terminal uint8[C], offsets int64[B+1], actions int16[C])."""
    actions, n = gl.legal_actions_batch(packed)
    n = n.to(torch.int64)
    offsets = torch.zeros((packed.shape[0] + 1,), dtype=torch.int64, device=packed.device)
    torch.cumsum(n, 0, out=offsets[1:])
    valid = torch.arange(gl.MAX_LEGAL, device=packed.device).unsqueeze(0) < n.unsqueeze(1)
    flat_actions = actions[valid].contiguous()
    parents = torch.repeat_interleave(packed, n, dim=0).contiguous()
    children, term = gl.next_batch(parents, flat_actions)
    return children, term, offsets, flat_actions


def _terminal_fixed(packed_terminal_flags):
    """terminal flags (bit0 is_lose, bit1 is_draw) -> int32 fixed values: -48 / 0 / NOT_FIXED (agents.py:69-73)."""
    t = packed_terminal_flags.to(torch.int32)
    fixed = torch.full_like(t, _NOT_FIXED)
    fixed = torch.where((t & 2) != 0, torch.zeros_like(t), fixed)
    fixed = torch.where((t & 1) != 0, torch.full_like(t, -MAX_DIST_FROM_GOAL), fixed)
    return fixed


@torch.no_grad()
def negamax_batch(packed, max_depth=2, max_level_states=48_000_000):
    """Exact depth-limited negamax (the value alpha_beta_action's root loop compares, agents.py:101) for every root.

    -> dict(action int16[B] = alpha_beta_action(root, max_depth), value48 int32[B] = 48 x the best score,
            scores48 int32[sum n] root-child scores in legal_actions() order, offsets int64[B+1]).
    The tree has max_depth + 1 plies below the root; level sizes grow by up to 133 per ply, so callers with many
    roots and max_depth >= 2 should chunk (max_level_states bounds one level, ~1.5 GB at the default).
    """
    _lib.require_cuda(packed, "packed")
    dev = packed.device
    L = _lib.load()
    levels = []  # per level: (offsets, fixed values of the CHILDREN, actions)
    cur, cur_alive = packed, None
    root_children = None
    for ply in range(max_depth + 1):
        # terminal nodes are not expanded; give them zero children by expanding only the live rows
        if cur_alive is None:
            live_idx = None
            children, term, off_live, acts = _expand(cur)
            offsets = off_live
        else:
            live_idx = torch.nonzero(cur_alive).squeeze(1)
            children, term, off_live, acts = _expand(cur[live_idx].contiguous())
            cnt = torch.zeros((cur.shape[0],), dtype=torch.int64, device=dev)
            cnt[live_idx] = off_live[1:] - off_live[:-1]
            offsets = torch.zeros((cur.shape[0] + 1,), dtype=torch.int64, device=dev)
            torch.cumsum(cnt, 0, out=offsets[1:])
        if children.shape[0] > max_level_states:
            raise ValueError(f"negamax level with {children.shape[0]} states exceeds max_level_states; use fewer roots per call")
        fixed = _terminal_fixed(term)
        levels.append((offsets, fixed, acts))
        if ply == 0:
            root_children = (offsets, acts)
        cur, cur_alive = children, fixed == _NOT_FIXED
    # depth-0 leaves: heuristic unless terminal (agents.py:69-75); aq_shortest_paths applies the same precedence
    with torch.cuda.device(dev):
        value = shortest_paths_batch(cur, want_heuristic=False, want_leaf=True)["leaf48"] if cur.shape[0] else \
            torch.empty((0,), dtype=torch.int32, device=dev)
        best = None
        for ply in range(max_depth, -1, -1):
            offsets, _, _ = levels[ply]
            P = offsets.shape[0] - 1
            parent_fixed = levels[ply - 1][1] if ply > 0 else None
            out = torch.empty((P,), dtype=torch.int32, device=dev)
            best = torch.empty((P,), dtype=torch.int32, device=dev)
            _lib.check(L.aq_negamax_backup(_lib.ptr(value), _lib.ptr(offsets), P, _lib.ptr(parent_fixed), _lib.ptr(out),
                                           _lib.ptr(best), _lib.stream_ptr(dev)), "aq_negamax_backup")
            if ply == 0:
                scores48 = -value
            value = out
    offsets, acts = root_children
    has = best >= 0
    idx = (offsets[:-1] + best.clamp(min=0).to(torch.int64)).clamp(max=max(acts.shape[0] - 1, 0))
    action = torch.where(has, acts[idx] if acts.shape[0] else torch.zeros_like(best, dtype=torch.int16),
                         torch.full((best.shape[0],), -1, dtype=torch.int16, device=dev))
    return {"action": action, "value48": value, "scores48": scores48, "offsets": offsets}


def alpha_beta_action(state, max_depth=2):
    """agents.py:90-107 for one State -> the action (None if the state has no legal action)."""
    a = int(negamax_batch(state._packed(), max_depth=max_depth)["action"][0].item())
    return None if a < 0 else a
