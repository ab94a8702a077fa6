import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
if GOLDEN_DIR not in sys.path:
    sys.path.insert(0, GOLDEN_DIR)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    arrays = np.load(os.path.join(GOLDEN_DIR, "golden.npz"))
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as fh:
        meta = json.load(fh)
    return arrays, meta


@pytest.fixture(scope="session")
def shipped(golden):
    from chromevol_enhanced_b200.newick import Tree
    _, meta = golden
    return meta["shipped_newick"], dict(meta["shipped_counts"])


def make_tree(newick):
    from chromevol_enhanced_b200.newick import Tree
    return Tree(newick, format=1)
