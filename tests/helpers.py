"""Shared test utilities: golden fixtures, graph reconstruction, comparison metrics, simulator engine."""
import os
import subprocess

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CSRC = os.path.join(ROOT, "connectome-gnn-suite_b200", "csrc")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "_build", "libcgnn_emu.so")
ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "liboracle.so")

# fp32 tolerance of BASELINE.json's north star: per-tensor max-norm relative error <= 1e-5
# (max|a-b| / max|ref|), evaluated with dropout disabled.  See SURVEY.md 8a note / A.3.
REL_TOL = 1e-5


def golden(name: str) -> dict:
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def graphs_from_store(a: dict) -> list:
    """Rebuild the list of ConnectomeGraph a fixture was made from (package class, CPU tensors)."""
    from connectome_gnn.graph import ConnectomeGraph
    out = []
    npz, epz = a["store.node_ptr"], a["store.edge_ptr"]
    for s in range(len(npz) - 1):
        n0, n1, e0, e1 = npz[s], npz[s + 1], epz[s], epz[s + 1]
        ei = torch.from_numpy(np.stack([a["store.src"][e0:e1], a["store.dst"][e0:e1]]).astype(np.int64))
        label = torch.tensor(int(a["store.label"][s]), dtype=torch.long) if a["store.has_label"][s] else None
        out.append(ConnectomeGraph(torch.from_numpy(a["store.x"][n0:n1].copy()), ei,
                                   torch.from_numpy(a["store.w"][e0:e1].copy()), label, f"sub-{s:04d}"))
    return out


def state_dict_from(a: dict, prefix: str) -> dict:
    """Tensors stored under '<prefix>.<state_dict key>'."""
    pre = prefix + "."
    return {k[len(pre):]: torch.from_numpy(np.array(v)) for k, v in a.items() if k.startswith(pre)}


def max_rel(got, ref) -> float:
    """Per-tensor max-norm relative error: max|got - ref| / max|ref| (0/0 -> 0)."""
    got = torch.as_tensor(got).detach().double().cpu().reshape(-1)
    ref = torch.as_tensor(ref).detach().double().cpu().reshape(-1)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if ref.numel() == 0:
        return 0.0
    denom = float(ref.abs().max())
    err = float((got - ref).abs().max())
    return err / denom if denom > 0 else err


def assert_close(got, ref, what: str, tol: float = REL_TOL, atol: float = 0.0):
    got_t, ref_t = torch.as_tensor(got).detach().cpu(), torch.as_tensor(ref).detach().cpu()
    assert torch.isfinite(got_t.double()).all(), f"{what}: non-finite values"
    if atol > 0.0:
        err = float((got_t.double() - ref_t.double()).abs().max()) if ref_t.numel() else 0.0
        assert err <= atol or max_rel(got_t, ref_t) <= tol, f"{what}: abs err {err:.3e} > {atol:.1e}"
        return
    r = max_rel(got_t, ref_t)
    assert r <= tol, f"{what}: max-norm relative error {r:.3e} > {tol:.1e}"


def build_emu() -> str:
    subprocess.run(["make", "-s", "-j8", "-C", CSRC, "emu"], check=True)
    return EMU_LIB


def build_oracle() -> str:
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    return ORACLE_LIB


_EMU = None


def emu_engine():
    """Engine over tests/emu/_build/libcgnn_emu.so: the kernel sources compiled for the single-threaded
    simulator, called with host pointers.  Test-only; the package never constructs this."""
    global _EMU
    if _EMU is None:
        from connectome_gnn import _engine, _lib

        class EmuEngine(_engine.Engine):
            def stream(self) -> int:
                return 0

        _EMU = EmuEngine(_lib.bind(build_emu()), torch.device("cpu"))
    return _EMU
