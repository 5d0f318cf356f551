"""The drop-in boundary: lib/libcgnn.so loads without a GPU and exports every symbol include/cgnn.h declares;
the ctypes prototypes of the host package cover the same set.  No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

import helpers

HEADER = os.path.join(helpers.ROOT, "include", "cgnn.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cgnn_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def product_lib():
    from connectome_gnn import _lib
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.run(["make", "-s", "-j8", "-C", helpers.CSRC, "lib"], check=True)
    return _lib.LIB_PATH


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for required in ("cgnn_collate_csr", "cgnn_csr_from_coo", "cgnn_gcn_layer_fwd", "cgnn_gcn_layer_bwd",
                     "cgnn_sage_layer_fwd", "cgnn_sage_layer_bwd", "cgnn_bn_finalize", "cgnn_bn_eval_affine",
                     "cgnn_bn_merge_stats", "cgnn_bn_bwd_sums", "cgnn_pool_fwd", "cgnn_head_fwd", "cgnn_head_bwd",
                     "cgnn_ce_fwd", "cgnn_ce_bwd"):
        assert required in names


def test_library_exports_every_declared_symbol(product_lib):
    lib = ctypes.CDLL(product_lib)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} is declared in include/cgnn.h but not exported by libcgnn.so"
    lib.cgnn_abi_version.restype = ctypes.c_int
    assert lib.cgnn_abi_version() == 11
    lib.cgnn_status_string.restype = ctypes.c_char_p
    assert lib.cgnn_status_string(0) == b"ok" and b"shared memory" in lib.cgnn_status_string(2)


def test_python_prototypes_cover_the_header(product_lib):
    from connectome_gnn import _lib
    assert sorted(_lib.PROTOTYPES) == declared_functions()
    bound = _lib.bind(product_lib)
    assert bound.cgnn_workspace_bytes() >= 1 << 20


def test_library_is_sm100a_native(product_lib):
    """The shipped code is sm_100a SASS (no PTX-JIT for another arch, no multi-arch fat binary)."""
    out = subprocess.run(["cuobjdump", "--list-elf", product_lib], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from connectome_gnn import _lib
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libcgnn.so"))
    with pytest.raises(_lib.CgnnError, match="no CPU"):
        _lib.load()


def test_no_cpu_path_without_cuda():
    """Product entry points refuse host tensors instead of computing on the CPU."""
    import torch
    from connectome_gnn import _engine
    with pytest.raises(RuntimeError, match="no CPU"):
        _engine.engine_for(torch.zeros(3))
    if not torch.cuda.is_available():
        from connectome_gnn import collate_graphs, generate_dataset
        with pytest.raises(RuntimeError, match="no CPU"):
            collate_graphs(generate_dataset(2, 10, seed=0))


def test_product_package_never_imports_the_oracle_or_simulator():
    pkg = os.path.join(helpers.ROOT, "connectome-gnn-suite_b200", "connectome_gnn")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("oracle fixture", ""), fn
            assert "cgnn_emu" not in src and "libcgnn_emu" not in src, fn
