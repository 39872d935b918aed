"""CPU: the C-ABI library builds/loads without a GPU and exports exactly what include/sais_b200.h declares."""
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "sais_b200.h"


def _declared():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(sais_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    names = _declared()
    for must in ("sais_gemm_bias_act", "sais_layernorm", "sais_vit_attention", "sais_normalize_patchify_u8",
                 "sais_temporal_prep", "sais_temporal_attention", "sais_clip_head", "sais_prototype_score",
                 "sais_vit_forward", "sais_temporal_forward", "sais_version", "sais_last_error"):
        assert must in names


def test_library_loads_and_exports_every_declared_symbol():
    from sais_b200 import _lib
    handle = _lib.lib()  # builds in-tree if missing; raises if it cannot
    declared = _declared()
    assert sorted(_lib.SIGNATURES) == declared, "ctypes table and header disagree"
    for name in declared:
        assert hasattr(handle, name), f"{name} not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (sais_[a-z0-9_]+)\b", out))
    assert exported == set(declared)
    assert handle.sais_version() >= 100
    assert handle.sais_launch_count() == 0 or handle.sais_launch_count() > 0  # callable without a device


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors of the header structs: pointer-count bookkeeping catches a forgotten field."""
    import ctypes as C
    from sais_b200 import _lib
    assert C.sizeof(_lib.SaisVitBlockWeights) == 18 * 8
    assert C.sizeof(_lib.SaisVitWeights) == (4 + 12 * 18 + 2) * 8
    assert C.sizeof(_lib.SaisTemporalLayerWeights) == 12 * 8
    assert C.sizeof(_lib.SaisTemporalWeights) == 2 * 8 + 8 + 4 * 12 * 8  # n_pos int32 padded to 8
    assert C.sizeof(_lib.SaisGemmArgs) == 7 * 8 + 8 * 8 + 4 * 4 + 5 * 8 + 2 * 4


def test_no_cpu_fallback():
    """product modules refuse CPU tensors instead of silently computing on the host."""
    import torch
    from sais_b200 import SaisError, ops
    with pytest.raises(SaisError):
        ops.layernorm(torch.zeros(4, 384), torch.ones(384), torch.zeros(384), 1e-6)
    with pytest.raises(SaisError):
        ops.prototype_score(torch.zeros(2, 256), torch.zeros(2, 256))
    with pytest.raises(SaisError):  # fused MLP with the folded LayerNorm (the default MLP path of sais_vit_forward)
        ops.vit_mlp_ln(torch.zeros(4, 384, dtype=torch.bfloat16), torch.zeros(4, 8), torch.zeros(1536, 384, dtype=torch.bfloat16),
                       torch.zeros(1536), torch.zeros(1536), torch.zeros(384, 1536, dtype=torch.bfloat16), torch.zeros(384),
                       torch.zeros(4, 384))
    with pytest.raises(SaisError):
        ops.rowstats_cast(torch.zeros(4, 384))


def test_product_never_imports_oracle():
    for py in (ROOT / "sais_b200").rglob("*.py"):
        assert "oracle" not in py.read_text(), f"{py} references the oracle"
    for cu in (ROOT / "sais_b200" / "csrc").glob("*.cu*"):
        assert "oracle" not in cu.read_text()


def test_hot_kernels_do_not_spill():
    """Register-budget regression guard (ptxas -v logs written by the build): the 16-warp GELU epilogue only pays off without
    spills — earlier 12 / 16-warp variants were capped at 128 / 96 registers and lost 20-60 us to local-memory traffic
    (DESIGN.md 3.2).  A few bytes for loop-invariant values are tolerated; the fused MLP and the 8-warp kernels must be clean."""
    import re
    from pathlib import Path
    build = Path(__file__).resolve().parents[1] / "sais_b200" / "csrc" / "build"
    logs = {n: build / f"{n}.ptxas.log" for n in ("gemm_tcgen05", "mlp_fused")}
    if not all(p.exists() for p in logs.values()):
        pytest.skip("no ptxas logs (library not built in-tree)")

    def spills(text):
        out = {}
        for m in re.finditer(r"Compiling entry function '(\S+)'.*?(\d+) bytes spill stores, (\d+) bytes spill loads", text, re.S):
            out[m.group(1)] = (int(m.group(2)), int(m.group(3)))
        return out

    gemm = spills(logs["gemm_tcgen05"].read_text())
    assert gemm, "no kernels found in the ptxas log"
    lean = {k: v for k, v in gemm.items() if re.search(r"ELi16ELb[01]E", k)}   # EW = 16 instantiations
    eight = {k: v for k, v in gemm.items() if re.search(r"ELi8ELb[01]E", k)}   # EW = 8 instantiations
    assert lean and eight
    assert max(v[0] for v in lean.values()) <= 64, {k[-40:]: v for k, v in lean.items() if v[0] > 64}
    assert max(v[0] for v in eight.values()) <= 64, {k[-40:]: v for k, v in eight.items() if v[0] > 64}
    mlp = spills(logs["mlp_fused"].read_text())
    assert mlp and all(v == (0, 0) for v in mlp.values()), mlp
