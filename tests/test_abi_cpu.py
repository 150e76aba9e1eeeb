"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol the header
declares (no compute without a GPU), the ctypes structs match the header, and the product path
fails loudly without a CUDA device instead of falling back to anything on the CPU."""
import ctypes as C
import os
import re

import pytest

from conftest import REPO


def _build():
    import __graft_entry__
    __graft_entry__.build()


def test_library_exports_every_declared_symbol():
    _build()
    from gym_cellular_b200 import _lib
    header = open(os.path.join(REPO, "include", "gym_cellular_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(gc_[a-z_]+)\s*\(", header)))
    assert declared == sorted(_lib.EXPORTS)
    lib = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().gc_abi_version() == _lib.ABI_VERSION == int(re.search(r"GC_ABI_VERSION (\d+)", header).group(1))


def test_struct_layout_matches_header():
    from gym_cellular_b200 import _lib
    # gc_config: 8 x 32-bit, then 3 x int64, uint64, 2 x double -> 80 bytes, no padding
    assert C.sizeof(_lib.GcConfig) == 8 * 4 + 4 * 8 + 2 * 8 == 80
    assert _lib.GcConfig.n_envs.offset == 32 and _lib.GcConfig.seed.offset == 56
    assert C.sizeof(_lib.GcCellTables) == 9 * C.sizeof(C.c_void_p)
    header = open(os.path.join(REPO, "include", "gym_cellular_b200.h")).read()
    for name, val in (("GC_F_NOISE", _lib.F_NOISE), ("GC_F_RNG_EPISODIC", _lib.F_RNG_EPISODIC),
                      ("GC_F_REWARD_LOG2", _lib.F_REWARD_LOG2), ("GC_F_GENERIC_KERNEL", _lib.F_GENERIC_KERNEL), ("GC_MAX_CELLS", _lib.MAX_CELLS),
                      ("GC_MAX_LEVELS", _lib.MAX_LEVELS), ("GC_N_STATS", _lib.N_STATS)):
        assert int(re.search(rf"#define {name}\s+(\d+)", header).group(1)) == val


def test_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call, so it is testable on the CPU box."""
    _build()
    from gym_cellular_b200 import _lib
    L = _lib.load()
    cfg = _lib.GcConfig()
    h = C.c_void_p()
    assert L.gc_create(C.byref(cfg), C.byref(h)) == _lib.ERR_INVALID
    assert b"struct_size" in L.gc_last_error()
    cfg.struct_size = C.sizeof(cfg)
    cfg.n_envs, cfg.ld, cfg.n_cells, cfg.n_states, cfg.n_actions = 10, 10, 3, 3, 3
    assert L.gc_create(C.byref(cfg), C.byref(h)) == _lib.ERR_INVALID and b"multiple of 16" in L.gc_last_error()
    cfg.ld, cfg.n_cells, cfg.n_states = 16, 16, 8
    assert L.gc_create(C.byref(cfg), C.byref(h)) == _lib.ERR_INVALID and b"32-bit" in L.gc_last_error()
    assert L.gc_step(None, 0, 0, *([None] * 13)) == _lib.ERR_INVALID


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import gym_cellular_b200 as pkg
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.CellularVectorEnv(num_envs=4)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under gym_cellular_b200/ may import, include, link or
    load it (comments may mention it)."""
    pkg = os.path.join(REPO, "gym_cellular_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|#\s*include\s*[\"<][^\">]*oracle|libgc_oracle|gco_[a-z_]+\s*\(", re.M)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                src = open(os.path.join(root, f)).read()
                assert not bad.search(src), os.path.join(root, f)
