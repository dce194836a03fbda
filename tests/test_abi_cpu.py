"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/rag_b200.h
declares; host-side module hygiene; argument errors are reported without touching a GPU."""
import copy
import os
import pickle
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "rag_b200.h")).read()
    return sorted(set(re.findall(r"RAG_API\s+[\w\s\*]+?\b(rag_\w+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    from rag_b200 import _cabi

    lib = _cabi.lib()
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rag_b200.h but not exported"
        assert n in _cabi.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_cabi.SIGNATURES) == names
    assert lib.rag_abi_version() == _cabi.ABI_VERSION
    m = re.search(r"#define RAG_B200_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "rag_b200.h")).read())
    assert int(m.group(1)) == _cabi.ABI_VERSION


def test_argument_errors_without_gpu():
    from rag_b200 import _cabi

    lib = _cabi.lib()
    assert lib.rag_cost_volume_fwd(None, None, None, 1, 1, 1, 1, 1, None) == -1
    assert b"null" in lib.rag_last_error()
    assert lib.rag_disp_head_fwd(None, None, None, 1, 1, 1, 1, 3, None) == -1
    assert lib.rag_loss_metrics_scratch(480, 960) > 0
    assert lib.rag_cv_stem_moments(None, None, None, None, 1, 12, 12, 8, 4, 8, None, None) == -1       # null pointers
    assert lib.rag_cost_volume_fwd_ws(None, None, None, 1, 1, 1, 1, 1, None, None) == -1
    assert lib.rag_cv_stem_workspace_bytes(12, 12) == 12 * 2 * 12 * 3 * 12 * 4
    assert lib.rag_disp_head_bwd(None, None, None, None, None, None, 1, 1, 1, 1, 3, None) == -1
    with pytest.raises(RuntimeError, match="code -1"):
        _cabi.check(-1, "x")


def test_sass_is_sm100a_only():
    import shutil
    import subprocess

    from rag_b200 import _cabi

    _cabi.lib()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "--list-elf", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_modules_are_stateless_and_cpu_inputs_raise():
    from rag_b200.modules import CostVolume, Disp, DisparityRegression

    for m in (CostVolume(192), Disp(192), DisparityRegression(192)):
        assert len(m.state_dict()) == 0 and not list(m.parameters()) and not list(m.buffers())
        m2 = pickle.loads(pickle.dumps(copy.deepcopy(m)))
        assert type(m2) is type(m) and m2.maxdisp == 192
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CostVolume(192)(torch.zeros(1, 2, 3, 4), torch.zeros(1, 2, 3, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Disp(192)(torch.zeros(1, 1, 64, 2, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DisparityRegression(192)(torch.zeros(1, 192, 2, 2))
    from rag_b200.fused_stem import cv_stem_batch_stats, cv_stem_forward
    from rag_b200.last_conv import conv3d_c1_forward

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cv_stem_forward(torch.zeros(1, 12, 3, 8), torch.zeros(1, 12, 3, 8), torch.zeros(12, 24, 3, 3, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cv_stem_batch_stats(torch.zeros(1, 12, 3, 8), torch.zeros(1, 12, 3, 8), torch.zeros(12, 24, 3, 3, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        conv3d_c1_forward(torch.zeros(1, 12, 3, 3, 8), torch.zeros(1, 12, 3, 3, 3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "rag_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(import\s+oracle|from\s+oracle|oracle[./]rag_oracle|oracle/)", src), f"{f} uses the oracle"
                assert "/root/reference" not in src
