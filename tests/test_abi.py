"""CPU tests of the C-ABI boundary: the library builds/loads, exports every symbol include/gta_b200.h declares,
and validates its arguments (no CUDA work is launched here)."""
import ctypes
import os
import re

import pytest

from gta_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "gta_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gta_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    decl = _declared_symbols()
    assert decl, "no declarations parsed"
    assert sorted(_lib.SYMBOLS) == decl


def test_flag_constants_match_header():
    """The GTA_FLAG_* / GTA_ERR_* / GTA_DTYPE_* values of the ctypes binding are the header's."""
    txt = open(os.path.join(ROOT, "include", "gta_b200.h")).read()
    defs = dict(re.findall(r"#define\s+(GTA_(?:FLAG|ERR|DTYPE)_[A-Z0-9_]+)\s+\(?(-?\d+)\)?", txt))
    assert "GTA_FLAG_BWD_SPLIT" in defs and "GTA_FLAG_SINGLE_LAUNCH" in defs
    for name, val in defs.items():
        if hasattr(_lib, name):
            assert getattr(_lib, name) == int(val), name
    for name in dir(_lib):
        if name.startswith("GTA_FLAG_"):
            assert name in defs, name


def test_library_exports_every_declared_symbol():
    l = _lib.lib()
    for name in _declared_symbols():
        assert hasattr(l, name), name
    assert l.gta_abi_version() == 5


def test_struct_layout_matches_header_order():
    names = [f[0] for f in _lib.GtaAttnParams._fields_]
    txt = open(os.path.join(ROOT, "include", "gta_b200.h")).read()
    body = txt[txt.index("typedef struct GtaAttnParams"):txt.index("} GtaAttnParams;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    order = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(const\s+)?(void|float|int64_t|int|size_t|GtaReps|long long)\s*\*?\s*", "", decl)
        order += [d.strip().lstrip("*") for d in decl.split(",")]
    assert order == names


def _header_fields(struct):
    txt = open(os.path.join(ROOT, "include", "gta_b200.h")).read()
    body = txt[txt.index("typedef struct " + struct):txt.index("} " + struct + ";")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    order = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(const\s+)?(void|float|int64_t|int|size_t|GtaReps|GtaAttnParams|long long)\s*\*?\s*", "", decl)
        order += [d.strip().lstrip("*") for d in decl.split(",")]
    return order


def test_reps_and_backward_struct_layouts_match_header_order():
    assert _header_fields("GtaReps") == [f[0] for f in _lib.GtaReps._fields_]
    assert _header_fields("GtaAttnBwdParams") == [f[0] for f in _lib.GtaAttnBwdParams._fields_]
    assert ctypes.sizeof(_lib.GtaAttnBwdParams) == ctypes.sizeof(_lib.GtaAttnParams) + 7 * 8


def test_backward_and_probs_argument_validation():
    l = _lib.lib()
    bp = _lib.GtaAttnBwdParams()
    bp.fwd = _params()
    assert l.gta_attn_bwd(ctypes.byref(bp), None) == -1 and "lse" in l.gta_last_error().decode()
    bp.fwd = _params(se3=12, so3=8, so2=6, t2=6)
    bp.fwd.reps = _lib.GtaReps(*([0x1000] * 6), None, 0x1000, 0x1000)
    # generic-path layouts (t2 block / unaligned blocks) have a backward of their own; it validates its arguments first
    assert l.gta_attn_bwd(ctypes.byref(bp), None) == -1 and "lse" in l.gta_last_error().decode()
    assert l.gta_attn_bwd_workspace_bytes_p(ctypes.byref(bp.fwd)) == 7 * 2048 + l.gta_attn_bwd_workspace_bytes(1, 2, 16, 16, 32)
    eu = _params(se3=15, so3=8, so2=0, euclid=1, D=96, triv=73)          # euclid_sim (padded head dim 128): served as well
    eu.reps = _lib.GtaReps(*([0x1000] * 7), None, None)
    bp.fwd = eu
    assert l.gta_attn_bwd(ctypes.byref(bp), None) == -1 and "lse" in l.gta_last_error().decode()
    dense = lambda T: (2 * T * 128 * 2 + 1023) // 1024 * 1024
    assert l.gta_attn_bwd_workspace_bytes_p(ctypes.byref(eu)) == 9 * dense(16) + l.gta_attn_bwd_workspace_bytes(1, 2, 16, 16, 128)
    assert l.gta_attn_bwd(None, None) == -1
    # Q' and dO' images, K' | V' images, delta, and (head dims <= 96: the fused kernel) the fp32 dQ' accumulation tiles
    assert l.gta_attn_bwd_workspace_bytes(1, 2, 16, 16, 32) == 2 * 2 * 8192 + 2 * 2 * 8192 + 1024 + 2 * 128 * 32 * 4
    assert l.gta_attn_bwd_workspace_bytes(1, 2, 16, 16, 128) == 2 * 2 * 32768 + 2 * 2 * 32768 + 1024
    p = _params()
    assert l.gta_attn_probs(ctypes.byref(p), None, None) == -1 and "lse" in l.gta_last_error().decode()
    assert l.gta_attn_probs_workspace_bytes(1, 2, 16, 16, 32) == 2 * 4096


def test_workspace_bytes():
    l = _lib.lib()
    # 2 tensors x B*H x ceil(Tk/128) tiles x 128*D*2 bytes + one ready flag (int) per tile for the single-launch kernel,
    # rounded up to 1 KiB
    flags = lambda units: (units * 4 + 1023) // 1024 * 1024
    assert l.gta_attn_fwd_workspace_bytes(2, 8, 1280, 96) == 2 * 2 * 8 * 10 * 128 * 96 * 2 + flags(2 * 8 * 10)
    assert l.gta_attn_fwd_workspace_bytes(1, 6, 600, 64) == 2 * 6 * 5 * 128 * 64 * 2 + flags(6 * 5)
    assert l.gta_attn_fwd_workspace_bytes(0, 6, 600, 64) == 0
    p = _params(B=2, H=8, Tk=1280, Tq=1280, D=96, se3=48, so3=24, so2=24)
    assert l.gta_attn_fwd_workspace_bytes_p(ctypes.byref(p)) == l.gta_attn_fwd_workspace_bytes(2, 8, 1280, 96)
    # generic path (t2 block): dense Q', K', V' (bf16), fp32 O', then the tile images
    g = _params(B=1, H=6, Tq=600, Tk=600, D=64, triv=2, se3=32, so3=0, so2=0, t2=30)
    dense = 3 * ((6 * 600 * 64 * 2 + 1023) // 1024 * 1024) + (6 * 600 * 64 * 4 + 1023) // 1024 * 1024
    assert l.gta_attn_fwd_workspace_bytes_p(ctypes.byref(g)) == dense + l.gta_attn_fwd_workspace_bytes(1, 6, 600, 64)


def test_pipeline_selection():
    """gta_attn_fwd_pipeline: the shape rule and the forcing flags (no CUDA work)."""
    l = _lib.lib()
    msn_enc = _params(B=64, H=8, Tk=1280, Tq=1280, D=96, se3=48, so3=24, so2=24)
    msn_dec = _params(B=64, H=8, Tk=1280, Tq=2560, D=96, se3=48, so3=24, so2=24)
    clevr = _params(B=32, H=6, Tk=600, Tq=600, D=64, se3=32, so3=0, so2=32)
    assert l.gta_attn_fwd_pipeline(ctypes.byref(msn_enc)) == 0      # large and rotation-heavy: staging kernel + attention kernel
    assert l.gta_attn_fwd_pipeline(ctypes.byref(msn_dec)) == 1      # one key tile to rotate per work item: one launch
    assert l.gta_attn_fwd_pipeline(ctypes.byref(clevr)) == 1        # small: one launch
    msn_enc.flags = _lib.GTA_FLAG_SINGLE_LAUNCH
    assert l.gta_attn_fwd_pipeline(ctypes.byref(msn_enc)) == 1
    clevr.flags = _lib.GTA_FLAG_TWO_LAUNCH
    assert l.gta_attn_fwd_pipeline(ctypes.byref(clevr)) == 0
    f32 = _params(B=2, H=6, Tk=600, Tq=600, D=64, se3=32, so3=0, so2=32, in_dtype=_lib.GTA_DTYPE_F32)
    assert l.gta_attn_fwd_pipeline(ctypes.byref(f32)) == 2
    f32.flags = _lib.GTA_FLAG_FAST_FP32
    assert l.gta_attn_fwd_pipeline(ctypes.byref(f32)) == 1
    t2 = _params(B=1, H=6, Tq=600, Tk=600, D=64, triv=2, se3=32, so3=0, so2=0, t2=30)
    assert l.gta_attn_fwd_pipeline(ctypes.byref(t2)) == 3


def test_dev_library_is_separate():
    """Probes, micro-benchmarks and the first-generation kernel live in libgta_b200_dev.so, not in the product ABI."""
    d = _lib.dev_lib()
    for name in _lib.DEV_SYMBOLS:
        assert hasattr(d, name), name
    prod = _lib.lib()
    for name in ("gta_umma_probe", "gta_umma_bench", "gta_softmax_bench", "gta_dev_umma_probe"):
        assert not hasattr(prod, name), name
    txt = open(os.path.join(ROOT, "include", "gta_b200_dev.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    assert sorted(set(re.findall(r"\b(gta_dev_[a-z0-9_]+)\s*\(", txt))) == sorted(_lib.DEV_SYMBOLS)


def _params(**over):
    p = _lib.GtaAttnParams()
    dummy = 0x1000
    p.q = p.k = p.v = p.out = dummy
    p.B, p.H, p.Tq, p.Tk, p.D = 1, 2, 16, 16, 32
    p.Nq = p.Nk = 2
    p.triv, p.se3, p.so3, p.so2 = 0, 16, 8, 8
    p.reps = _lib.GtaReps(dummy, dummy, dummy, dummy, dummy, dummy, None, None, None)
    p.q_stride_b = p.k_stride_b = p.v_stride_b = 16 * 2 * 32
    p.q_stride_h = p.k_stride_h = p.v_stride_h = 32
    p.q_stride_t = p.k_stride_t = p.v_stride_t = 64
    p.scale = 1.0
    p.v_transform = 1
    for k, v in over.items():
        setattr(p, k, v)
    return p


@pytest.mark.parametrize("over,code,msg", [
    (dict(D=48), -3, "head dim"),
    (dict(se3=20, so3=4), -1, "whole"),
    (dict(se3=18, so3=0, so2=14), -1, "whole"),
    (dict(se3=15, so3=8, so2=0, euclid=1, D=96, triv=73), -1, "se3_qi"),
    (dict(se3=12, so3=8, so2=6, t2=6), -1, "t2 coordinates"),
    (dict(D=128, se3=0, so3=0, so2=0, triv=128, euclid=1), -3, "euclid_sim"),
    (dict(se3=24), -1, "sum to head dim"),
    (dict(Tq=15), -1, "divisible"),
    (dict(q=None), -1, "null"),
    (dict(in_dtype=7), -3, "dtype"),
    (dict(q_stride_t=65), -3, "16-byte"),
    (dict(workspace=None), -1, "workspace"),
])
def test_argument_validation(over, code, msg):
    l = _lib.lib()
    p = _params(**over)
    rc = l.gta_attn_fwd(ctypes.byref(p), None)
    assert rc == code
    assert msg in l.gta_last_error().decode()
    assert l.gta_attn_fwd(None, None) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.GtaError, match="no CPU fallback"):
        _lib.lib()
