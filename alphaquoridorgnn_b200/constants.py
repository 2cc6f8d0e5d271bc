"""Board configuration (reference constants.py:1-23; the 9x9 block of constants.py:17-20, which is
the configuration BASELINE.json is quoted on).  The CUDA kernels are compiled for this board."""
BOARD_SIZE = 9
NUM_WALLS = 10
NUM_PLIES_FOR_DRAW = 116  # (10 wall placements + max 48 moves from goal) * 2

PV_NETWORK_NAME = 'GNN'  # which network to use
PV_NETWORK_PATH = f'models/{PV_NETWORK_NAME}/{BOARD_SIZE}x{BOARD_SIZE}/'  # path for network weights
