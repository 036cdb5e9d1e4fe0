"""CPU checks of the drop-in boundary: the C-ABI library loads and exports exactly the entry points
include/distbo_b200.h declares, the ctypes table mirrors the header, and the product refuses to run
without a CUDA device (no CPU fallback).  No compute call is made here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "distbo_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = re.findall(r"\b(?:int|int64_t)\s+(distbo_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S)
    out = {}
    for name, args in decls:
        args = args.strip()
        out[name] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    return out


def test_header_declares_the_path():
    fns = header_functions()
    for must in ["distbo_embed_fwd_f32", "distbo_embed_fwd_f64", "distbo_embed_bwd_f64", "distbo_task_gram_f64",
                 "distbo_gram_f64", "distbo_potrf_f64", "distbo_trtri_f64", "distbo_lauum_f64", "distbo_nll_f64",
                 "distbo_gram_grad_f64", "distbo_task_gram_bwd_f64", "distbo_adam_f64", "distbo_score_f64"]:
        assert must in fns, must


def test_library_exports_every_header_symbol():
    from distbo_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the library first (python -c 'import __graft_entry__ as g; g.build()')"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), "libdistbo_b200.so does not export %s" % name


def test_ctypes_table_mirrors_header():
    from distbo_b200 import _lib
    fns = header_functions()
    assert set(_lib.SIGNATURES) == set(fns), set(_lib.SIGNATURES) ^ set(fns)
    for name, nargs in fns.items():
        assert len(_lib.SIGNATURES[name][1]) == nargs, (name, nargs, len(_lib.SIGNATURES[name][1]))
    lib = _lib.load()
    assert lib.distbo_version() == 1000          # compiled for sm_100a
    assert lib.distbo_launch_count(0) == 0       # nothing launched by loading


def test_sass_is_sm100a_with_fp64_tensor_ops():
    """The shared library carries sm_100a SASS and the Cholesky / scoring kernels use DMMA."""
    import shutil
    import subprocess
    from distbo_b200 import _lib
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    elf = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    per_fn = sass.split("Function : ")[1:]
    for kernel in ("score_kernel", "gemm_tiles_kernel", "potf2_inv_kernel"):
        mangled = "%d%s" % (len(kernel), kernel)           # exact Itanium name component, not a substring
        bodies = [b for b in per_fn if mangled in b.split("\n", 1)[0]]
        assert bodies and all("DMMA" in b for b in bodies), kernel
    # K1 fp32 forward: rows arrive by bulk async copies (the 1-D TMA path, SASS UBLKCP) guarded by mbarriers (SYNCS),
    # layer 1 runs on the tensor pipe (mma.sync TF32 -> HMMA.1688.F32.TF32)
    bodies = [b for b in per_fn if "19embed_fwd_tc_kernel" in b.split("\n", 1)[0]]
    assert len(bodies) == 8            # column tiles 1..4 x row tiles 1..2
    for b in bodies:
        assert "UBLKCP" in b and "SYNCS" in b and "HMMA.1688.F32.TF32" in b
    # tcgen05 digit-plane contraction (scoring, LAUUM, top levels of TRTRI): TMA tensor loads (UTMALDG), int8 MMAs
    # with accumulators in tensor memory (UTCIMMA, the second MMA of a digit re-using the A tile from the
    # collector), tensor-memory read-back (LDTM), no legacy tensor instruction
    bodies = [b for b in per_fn if "15score_i8_kernel" in b.split("\n", 1)[0]]
    assert len(bodies) == 3            # <0> scoring epilogue (score.cu), <1> store epilogue (inverse_i8.cu, chol.cu)
    for b in bodies:
        assert b.count("UTCIMMA") == 20 and "A_KEEP" in b and "A_REUSE" in b
        assert b.count("UTMALDG.3D") == 3 and "LDTM" in b and "UTCBAR" in b
        assert "HMMA" not in b and "DMMA" not in b and "IMMA." not in b


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from distbo_b200 import _lib, engine
    with pytest.raises(_lib.DistboError):
        engine.GPEngine(2)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "distbo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)


def test_dimension_limits_are_reported():
    """Widths the kernels keep in registers are checked on the host with a message (the C ABI itself returns -1)."""
    from distbo_b200.engine import NetSpec
    NetSpec(166, 2, 20, 10, y_mode=2)
    NetSpec(7, 1, 32, 16, 32, 16, y_mode=3)
    with pytest.raises(ValueError):
        NetSpec(7, 1, 33, 10)
    with pytest.raises(ValueError):
        NetSpec(7, 1, 20, 10, 20, 17, y_mode=1)
    with pytest.raises(ValueError):
        NetSpec(7, 1, 20, 10, y_mode=5)


def test_digit_plane_planning_entry_points_are_pure_host():
    """Work-list sizes and level selection of the tcgen05 inverse products need no device (the plan functions that
    upload them do): sizes grow with n_pad, the int8 route of the triangular inverse starts where the halving tree's
    blocks reach min_rows, bad arguments are refused."""
    from distbo_b200 import _lib
    lib = _lib.load()
    assert lib.distbo_trtri_i8_first_level(8192, 1024) == 4      # heights 1..3 merge 128/256/512-row blocks
    assert lib.distbo_trtri_i8_first_level(8192, 128) == 1
    assert lib.distbo_trtri_i8_first_level(2560, 1024) == 5      # 20 blocks: 10 + 10 is the first split >= 1024 rows
    assert lib.distbo_trtri_i8_first_level(1000, 1024) == -1     # not a multiple of 128
    w1, w2 = lib.distbo_trtri_i8_sched_words(4096, 1024), lib.distbo_trtri_i8_sched_words(8192, 1024)
    assert 0 < w1 < w2
    assert lib.distbo_trtri_i8_sched_words(8192, 64) == 0
    l1, l2 = lib.distbo_lauum_i8_sched_words(2048), lib.distbo_lauum_i8_sched_words(8192)
    assert 0 < l1 < l2 and lib.distbo_lauum_i8_sched_words(100) == 0
    # items of LAUUM: sum_ib 2 (ib + 1) (128-row tile, 64-row block) pairs, 4 words each, + offsets
    nb = 8192 // 128
    assert l2 >= 4 * sum(2 * (ib + 1) for ib in range(nb))
    assert lib.distbo_lauum_i8_workspace(1024) == 1024 * 1024 + 1024
    assert lib.distbo_trtri_i8_workspace(1024) == 4 * 1024 * 1024 + 4 * 1024
    assert lib.distbo_score_i8_workspace(1024) > 2 * lib.distbo_score_chunk() * 1024
