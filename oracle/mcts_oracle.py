"""CPU restatement of the reference tree search -- TEST INFRASTRUCTURE / CPU BASELINE, NOT PRODUCT.

``pv_mcts_scores`` follows /root/reference/pv_mcts.py:20-95 literally (Node.evaluate: terminal values -1 / 0
(35-42), leaf expansion through ``model.predict`` with one child per legal action (45-57), negamax recursion
(60-66), PUCT ``(-w/n if n else 0) + 1.25 * p * sqrt(t) / (1 + n)`` with np.argmax's first-maximum tie-break
(69-78), visit counts of the root's children (88)).  The states it walks are ``COracleState``: the game rules
come from the C oracle (oracle/quoridor_oracle.c), itself pinned against the unmodified reference.

Pinned by tests/test_oracle_golden.py::test_mcts_port_matches_reference_visit_counts against
tests/golden/mcts_golden.json (visit counts of the UNMODIFIED reference ``pv_mcts_policy`` under deterministic
evaluators, 50 and 200 simulations).  bench.py times it, with the torch-CPU GNN oracle as ``model.predict``, as the
CPU baseline of the MCTS metric (kind "port").
"""
from math import sqrt

import numpy as np

from oracle import quoridor_oracle as qo

C_PUCT = 1.25  # pv_mcts.py:71


class COracleState:
    """game_logic.State surface used by the search (legal_actions, next, is_lose, is_draw, is_done), on a row68 + plies
    record with the C oracle behind it."""
    __slots__ = ("row", "plies", "_la")

    def __init__(self, row=None, plies=0):
        if row is None:
            row = np.zeros(68, np.uint8)
            row[0], row[1], row[2], row[3] = 76, 10, 76, 10  # game_logic.py:31-37 on the 9x9 board
        self.row = np.ascontiguousarray(row, dtype=np.uint8)
        self.plies = int(plies)
        self._la = None

    @property
    def plies_played(self):
        return self.plies

    def is_lose(self):  # game_logic.py:43-46
        return int(self.row[2]) // 9 == 0

    def is_draw(self):  # game_logic.py:49-50
        return self.plies >= 116

    def is_done(self):
        return self.is_lose() or self.is_draw()

    def legal_actions(self):
        if self._la is None:
            out = qo.legal_actions_batch(self.row[None, :], np.array([self.plies], np.int16), nthreads=1)
            self._la = out["actions"][0, : int(out["n"][0])].tolist()
        return self._la

    def next(self, action):
        rows, plies, _ = qo.next_batch(self.row[None, :], np.array([self.plies], np.int16), np.array([action], np.int16))
        return COracleState(rows[0], int(plies[0]))


class _Node:  # pv_mcts.py:24-78
    __slots__ = ("state", "p", "w", "n", "child_nodes")

    def __init__(self, state, p):
        self.state, self.p, self.w, self.n, self.child_nodes = state, p, 0, 0, None

    def evaluate(self, predict):
        if self.state.is_done():
            value = -1 if self.state.is_lose() else 0
            self.w += value
            self.n += 1
            return value
        if not self.child_nodes:
            prior, value = predict(self.state)
            self.w += value
            self.n += 1
            self.child_nodes = [_Node(self.state.next(a), p) for a, p in zip(self.state.legal_actions(), prior)]
            return value
        value = -self.next_child_node().evaluate(predict)
        self.w += value
        self.n += 1
        return value

    def next_child_node(self):
        t = sum(c.n for c in self.child_nodes)
        pucb = [(-c.w / c.n if c.n else 0.0) + C_PUCT * c.p * sqrt(t) / (1 + c.n) for c in self.child_nodes]
        return self.child_nodes[int(np.argmax(pucb))]


def pv_mcts_scores(predict, state, sims):
    """-> visit counts of the root's children in state.legal_actions() order (pv_mcts.py:81-88)."""
    root = _Node(state, 0)
    for _ in range(sims):
        root.evaluate(predict)
    return [c.n for c in root.child_nodes]
