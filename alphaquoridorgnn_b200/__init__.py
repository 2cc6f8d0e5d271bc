"""B200-native (sm_100a) implementation of the AlphaQuoridorGNN hot path: batched Quoridor legal
moves / wall legality, board-graph construction, the pv_network_gnn GCN policy-value network
(forward + backward), batched leaf evaluation and lock-step PV-MCTS.  See DESIGN.md."""
from . import _lib  # noqa: F401
from .constants import BOARD_SIZE, NUM_WALLS, NUM_PLIES_FOR_DRAW  # noqa: F401

__all__ = ["game_logic", "pv_network_gnn", "positions", "constants"]
