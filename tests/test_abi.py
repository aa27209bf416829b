"""The C-ABI library loads and exports every symbol include/agrl_b200.h declares; the ctypes table
matches the header; compute entry points fail loudly without a B200 (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'agrl_b200.h')).read()
    return sorted(set(re.findall(r'^AGRL_API[^;(]*?\b(agrl_\w+)\s*\(', text, flags=re.M)))


def test_header_declares_the_three_subsystems():
    names = declared_symbols()
    for must in ('agrl_rank_market1501_dev', 'agrl_rank_mars_dev', 'agrl_distance_dev', 'agrl_head_forward_dev',
                 'agrl_rank_market1501_host', 'agrl_rank_mars_host', 'agrl_distance_host'):
        assert must in names


def test_library_exports_every_declared_symbol():
    from agrl.pytorch_b200 import _lib
    lib = _lib.load()
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(raw, name), name
    assert sorted(_lib.exported_names()) == declared_symbols()
    assert lib.agrl_abi_version() == 2
    assert lib.agrl_status_string(0) == b'ok'
    assert b'no CPU fallback' in lib.agrl_status_string(_lib.E_NO_DEVICE)


def test_library_exports_nothing_but_the_header():
    """no diagnostic entry points (e.g. the -DAGRL_TIMELINE build's agrl_timeline_set) in the product library"""
    import subprocess
    from agrl.pytorch_b200 import _lib
    _lib.load()
    out = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(line.split()[-1] for line in out.splitlines() if ' T ' in line and line.split()[-1].startswith('agrl_'))
    assert exported == declared_symbols()


def test_workspace_queries_need_no_device():
    from agrl.pytorch_b200 import _lib
    lib = _lib.load()
    assert lib.agrl_rank_workspace_bytes(1980, 9330, 50) > 4 * (2 * 1980 + 2 * 9330)
    # 3 bf16 planes of both operands + norms
    assert lib.agrl_distance_workspace_bytes(1980, 9330, 2048, 3) >= 3 * 2 * 2048 * (1980 + 9330)
    # the default split of the distance matrix: two fp16 planes + norms + per-row scales
    assert lib.agrl_distance_workspace_bytes(1980, 9330, 2048, _lib.SPLIT_FP16X2) >= (2 * 2 * 2048 + 8) * (1980 + 9330)
    assert lib.agrl_distance_workspace_bytes(10, 10, 2048, 7) == 0
    P = _lib.HeadParams()
    P.channels, P.num_layers, P.use_pose, P.learn_graph, P.split = 2048, 2, 1, 1, 2
    assert lib.agrl_head_prepared_bytes(ctypes.byref(P)) >= 2 * 2 * 2 * 2048 * 2048
    assert lib.agrl_head_workspace_bytes(ctypes.byref(P), 64, 8) >= 2 * 4 * 64 * 56 * 2048
    P.channels = 100
    assert lib.agrl_head_prepared_bytes(ctypes.byref(P)) == 0


def _no_gpu():
    import torch
    return not torch.cuda.is_available()


@pytest.mark.skipif(not _no_gpu(), reason='checks the no-device behaviour')
def test_compute_calls_fail_loudly_without_a_gpu():
    import torch
    from agrl.pytorch_b200 import _lib, metrics
    assert _lib.load().agrl_device_ok() == _lib.E_NO_DEVICE
    d = np.zeros((2, 3), np.float32)
    ids = np.zeros(2, np.int64), np.zeros(3, np.int64), np.zeros(2, np.int64), np.ones(3, np.int64)
    with pytest.raises(_lib.AgrlError, match='no CPU fallback'):
        metrics.evaluate_rank(d, *ids, use_metric_market1501=True)
    with pytest.raises(_lib.AgrlError, match='no CPU fallback'):
        metrics.evaluate_rank(d, *ids, use_metric_mars=True, max_rank=2)
    with pytest.raises(_lib.AgrlError, match='no CPU fallback'):
        metrics.compute_distance_matrix(torch.zeros(2, 4), torch.zeros(3, 4))


def test_product_never_imports_the_oracle():
    """nothing under agrl/ may import, call, link or execute anything under oracle/"""
    pkg = os.path.join(ROOT, 'agrl')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
                assert 'oracle/' not in text and 'liboracle' not in text, f


def test_header_is_plain_c(tmp_path):
    """the boundary is a C ABI: the header must compile as C99 (no C++-isms, no torch / CUDA types in the signatures)"""
    import shutil
    import subprocess
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no gcc')
    src = tmp_path / 'abi.c'
    src.write_text('#include "agrl_b200.h"\nint main(void) { return agrl_abi_version() == 0; }\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include')
    r = subprocess.run([gcc, '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror', '-I', inc, '-c', str(src), '-o',
                        str(tmp_path / 'abi.o')], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
