"""Import the UNMODIFIED reference modules from /root/reference (fixture generation only).

This file is used only by ``make_golden.py`` in the build container, where
``/root/reference`` exists.  Nothing in the tests, ``smoke()`` or ``bench.py``
imports it at run time (the GPU box has no ``/root/reference``).

The reference does not import cleanly as checked in:
  * ``game_logic.py:9`` imports ``code_profiling_util`` which imports
    ``snakeviz.cli`` (not installed)            -> dummy ``snakeviz`` modules;
  * ``constants.py:13-15`` is set to 5x5 and ``NUM_PLIES_FOR_DRAW`` is bound by
    value at import (``game_logic.py:8``)       -> inject a 9x9 ``constants``
    module with the values of ``constants.py:17-20``;
  * ``pv_mcts.py:8,14`` imports ``pv_network_cnn`` (needs ``torchsummary``,
    ``torch_tensorrt``) and a non-existent ``train_network.preprocess_input``
                                                -> dummy modules.
"""
import importlib
import sys
import types

REFERENCE_ROOT = "/root/reference"


def _dummy(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_reference(board_size=9, num_walls=10, plies_for_draw=116):
    """Returns (game_logic, pv_mcts) reference modules configured for the given board."""
    for name in ("constants", "game_logic", "pv_mcts", "code_profiling_util",
                 "pv_network_cnn", "BaseNetwork", "train_network"):
        sys.modules.pop(name, None)
    _dummy("snakeviz")
    _dummy("snakeviz.cli", main=lambda *a, **k: None)
    _dummy("torchsummary", summary=lambda *a, **k: None)
    _dummy("torch_tensorrt")
    _dummy("constants", BOARD_SIZE=board_size, NUM_WALLS=num_walls,
           NUM_PLIES_FOR_DRAW=plies_for_draw, PV_NETWORK_NAME="CNN",
           PV_NETWORK_PATH=f"models/CNN/{board_size}x{board_size}/")
    _dummy("train_network", preprocess_input=None)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    game_logic = importlib.import_module("game_logic")
    pv_mcts = importlib.import_module("pv_mcts")
    return game_logic, pv_mcts
