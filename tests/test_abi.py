"""CPU checks of the C-ABI boundary: the library loads, exports every declared symbol, and the ctypes mirrors of the
parameter structs have the C compiler's size and field offsets.  No compute calls (no GPU here)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

from helpers import ROOT, pkg

HEADER = os.path.join(ROOT, "include", "jl_b200.h")


def test_library_loads_and_exports_every_declared_symbol():
    L = pkg()._lib
    lib = L.load()
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(jl_[a-z0-9_]+)\s*\(", text))
    assert declared, "no declarations parsed from the header"
    for name in declared:
        assert hasattr(lib, name), f"libjl_b200.so does not export {name}"
    assert declared == set(L.SYMBOLS), f"ctypes table and header differ: {declared ^ set(L.SYMBOLS)}"
    assert lib.jl_version() == 1


STRUCTS = {
    "jl_mel_cmvn_params": "MelCmvnParams",
    "jl_gemm_params": "GemmParams",
    "jl_layernorm_fwd_params": "LayerNormFwdParams",
    "jl_layernorm_bwd_params": "LayerNormBwdParams",
    "jl_attn_fwd_params": "AttnFwdParams",
    "jl_wfadapter_fwd_params": "WFAdapterFwdParams",
    "jl_attn_bwd_params": "AttnBwdParams",
    "jl_ctc_params": "CtcParams",
    "jl_ctc_greedy_params": "CtcGreedyParams",
    "jl_adamw_params": "AdamWParams",
    "jl_fusion_params": "FusionParams",
    "jl_colreduce_job": "ColReduceJob",
    "jl_wfadapter_pack_params": "WFAdapterPackParams",
    "jl_attadapter_fwd_params": "AttAdapterFwdParams",
    "jl_lnfold_pack_params": "LnFoldPackParams",
    "jl_lnproj_bwd_params": "LnProjBwdParams",
}


def test_ctypes_structs_match_the_c_layout(tmp_path):
    L = pkg()._lib
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for cname, pyname in STRUCTS.items():
        cls = getattr(L, pyname)
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)]).decode().splitlines()
    got = {}
    for ln in out:
        c, f, v = ln.split()
        got[(c, f)] = int(v)
    for cname, pyname in STRUCTS.items():
        cls = getattr(L, pyname)
        assert got[(cname, "size")] == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, f"{cname}.{fname}"


def test_missing_library_fails_loudly(monkeypatch):
    L = pkg()._lib
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", "/nonexistent/libjl_b200.so")
    with pytest.raises(ImportError):
        L.load()


def test_cpu_tensors_are_rejected():
    import torch
    ops = pkg().ops
    with pytest.raises(RuntimeError):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))


def test_comm_entry_points_reject_bad_arguments_without_a_gpu():
    """jl_comm_* argument checks (no device work): null handles / buffers give JL_EINVAL with a message; destroy(NULL) is OK."""
    P = pkg()
    L = P._lib
    lib = L.load()
    assert lib.jl_comm_destroy(None) == 0
    assert lib.jl_comm_allreduce(None, None, 16, None) == L.JL_EINVAL
    assert b"communicator" in lib.jl_last_error()
    assert lib.jl_comm_unique_id(None) == L.JL_EINVAL
    assert lib.jl_comm_rank(None, None, None) == L.JL_EINVAL
    with pytest.raises(ValueError):
        P.JLComm(b"short", 0, 1)
