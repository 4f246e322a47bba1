"""The C-ABI library loads on a CPU-only box and exports every symbol include/mova_b200.h declares (no compute)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "mova_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mova_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    from dualforce_b200 import _lib

    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 10
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in include/mova_b200.h but not exported by libmova_b200.so"
    assert set(_lib.SIGNATURES) <= set(syms), "python binding names a symbol the header does not declare"
    assert lib.mova_b200_abi_version() == _lib.ABI_VERSION


def test_no_cpu_fallback():
    """CPU tensors must be rejected loudly: the product path has no CPU implementation."""
    import torch

    import dualforce_b200 as B

    x = torch.zeros(4, 8, dtype=torch.bfloat16)
    with pytest.raises(B.MovaB200Error):
        B.ops.linear(x, x)
    with pytest.raises(B.MovaB200Error):
        B.ops.attention(x[None], x[None], x[None], 1)
    if not torch.cuda.is_available():
        with pytest.raises(B.MovaB200Error):
            B.install(object())


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dualforce_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "mova_oracle" not in src and "ref_loader" not in src, f"{fn} references the test oracle"
            assert "emulated_ops" not in src, f"{fn} references the tests' kernel emulation"


def test_every_ops_entry_point_rejects_cpu_tensors():
    """Every kernel front end of dualforce_b200.ops (old and new) raises on CPU tensors instead of computing anything."""
    import torch

    import dualforce_b200 as B

    bf = torch.zeros(8, 128, dtype=torch.bfloat16)
    f32 = torch.zeros(128, dtype=torch.float32)
    calls = [
        lambda: B.ops.layernorm(bf, 1e-6),
        lambda: B.ops.rmsnorm_rope_(bf, bf[0], 1e-6),
        lambda: B.ops.lse_merge(bf[None], torch.zeros(1, 1, 8), 1),
        lambda: B.ops.add_to_f32(bf),
        lambda: B.ops.patchify(torch.zeros(4, 2, 2, 2), (1, 2, 2)),
        lambda: B.ops.unpatchify(bf, (8,), (1,), 128),
        lambda: B.ops.sinusoidal_embedding(256, torch.zeros(1)),
        lambda: B.ops.gemv_f32(f32, bf),
        lambda: B.ops.cfg_euler_step(bf, None, torch.zeros(8, 128), 1.0, -0.1),
    ]
    for i, call in enumerate(calls):
        with pytest.raises(B.MovaB200Error):
            call()
    assert {n for n in B.ops.__all__ if not n.startswith(("EPI_", "ROPE_"))} >= {
        "linear", "attention", "layernorm", "rmsnorm_rope_", "lse_merge", "add_to_f32", "patchify", "unpatchify",
        "sinusoidal_embedding", "gemv_f32", "cfg_euler_step"}


def test_ctypes_signatures_match_the_header_argument_by_argument():
    """Every prototype of include/mova_b200.h against dualforce_b200._lib.SIGNATURES: same arity, and pointer /
    int / int64_t / float arguments bound as c_void_p / c_int / c_int64 / c_float in the same positions (a swapped
    int64 stride or float would otherwise only show as garbage on the device)."""
    import ctypes

    from dualforce_b200 import _lib

    text = open(os.path.join(ROOT, "include", "mova_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    seen = 0
    for m in re.finditer(r"\b(mova_b200_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        params = [] if args in ("void", "") else [a.strip() for a in args.split(",")]
        want = []
        for prm in params:
            if "*" in prm:
                want.append(ctypes.c_void_p)
            elif prm.startswith("int64_t"):
                want.append(ctypes.c_int64)
            elif prm.startswith("int"):
                want.append(ctypes.c_int)
            elif prm.startswith("float"):
                want.append(ctypes.c_float)
            else:
                raise AssertionError(f"{name}: unrecognised parameter type in '{prm}'")
        assert _lib.SIGNATURES[name][1] == want, f"{name}: header {params} vs ctypes {_lib.SIGNATURES[name][1]}"
        seen += 1
    assert seen == len(_lib.SIGNATURES)


def test_peer_entry_points_validate_arguments_before_touching_the_device():
    """csrc/peer.cu: bad arguments are rejected with a message (no CUDA call is made for them, so this runs on a
    CPU-only box): too many flags per push, non-positive epochs, unaligned or missing flag words."""
    import ctypes

    from dualforce_b200 import _lib

    lib = _lib.load()
    none = ctypes.c_void_p()
    arr = (ctypes.c_void_p * 1)()
    nb = (ctypes.c_int64 * 1)()
    assert lib.mova_b200_peer_push(0, arr, arr, nb, 33, arr, 0, 1, none, none) != 0
    assert "flags" in _lib.last_error()
    assert lib.mova_b200_peer_push(0, arr, arr, nb, 0, arr, 0, 0, none, none) != 0
    assert "epoch" in _lib.last_error()
    assert lib.mova_b200_peer_wait(none, 4, 1, 1000, 1, none) != 0
    assert lib.mova_b200_peer_wait(ctypes.c_void_p(12), 4, 1, 1000, 1, none) != 0  # not 8-byte aligned
    assert "aligned" in _lib.last_error()
    assert lib.mova_b200_peer_wait(ctypes.c_void_p(16), 4, 0, 1000, 1, none) != 0
    assert lib.mova_b200_peer_alloc(0, ctypes.byref(none), ctypes.create_string_buffer(64)) != 0
    assert lib.mova_b200_peer_open(None, ctypes.byref(none)) != 0
