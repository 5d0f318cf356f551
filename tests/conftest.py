"""pytest configuration: import paths, the ``gpu`` marker, shared fixtures.

``-m "not gpu"`` runs here (no GPU): oracle vs golden fixtures, host logic, C-ABI symbol export and
kernel-logic tests of the *same kernel sources* on the test-only simulator (tests/emu).
``-m gpu`` runs on a B200: the parity tests proper, through the product CUDA library.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "connectome-gnn-suite_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def emu_engine():
    """Engine bound to the simulator build of the kernels (host pointers, stream 0)."""
    import helpers
    return helpers.emu_engine()


@pytest.fixture()
def on_emu(monkeypatch, emu_engine):
    """Route the package's engine lookup to the simulator for the duration of one test."""
    import torch
    from connectome_gnn import _engine
    monkeypatch.setattr(_engine, "engine_for", lambda t: emu_engine)
    monkeypatch.setattr(_engine, "default_device", lambda: torch.device("cpu"))
    return emu_engine
