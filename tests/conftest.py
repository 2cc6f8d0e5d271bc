import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ka():
    with open(os.path.join(GOLDEN, "ka_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def traj():
    d = np.load(os.path.join(GOLDEN, "trajectories.npz"))
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def graph_golden():
    d = np.load(os.path.join(GOLDEN, "graph_golden.npz"))
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def mcts_golden():
    with open(os.path.join(GOLDEN, "mcts_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def agents_golden():
    with open(os.path.join(GOLDEN, "agents_golden.json")) as f:
        return json.load(f)
