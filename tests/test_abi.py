"""C-ABI checks that run WITHOUT a GPU: the library loads, exports every symbol include/gemb200.h
declares, and fails loudly (no CPU fallback) when no device is present."""
import ctypes
import os
import re

import pytest

from conftest import REPO


def header_symbols():
    txt = open(os.path.join(REPO, 'include', 'gemb200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(gemb_[a-z0-9_]+)\s*\(', txt)))


def test_exports_match_header(native_lib):
    from gem_b200 import _native
    syms = header_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(native_lib, s), 'libgemb200.so lacks %s' % s
    assert sorted(_native.EXPORTS) == syms
    assert native_lib.gemb_version() == 100


def test_struct_sizes_are_stable(native_lib):
    from gem_b200 import _native
    assert ctypes.sizeof(_native.HopeOpts) == 56
    assert ctypes.sizeof(_native.HopeStats) == 104
    assert ctypes.sizeof(_native.N2VStats) == 120


def test_no_cpu_fallback(native_lib):
    from gem_b200 import _native
    if native_lib.gemb_device_count() > 0:
        pytest.skip('a GPU is present')
    with pytest.raises(RuntimeError, match='no CUDA device'):
        _native.Context(0)
    import networkx as nx
    from gem_b200.embedding.hope import HOPE
    G = nx.DiGraph([(0, 1), (1, 2)])
    with pytest.raises(RuntimeError):
        HOPE(d=2, beta=0.01).learn_embedding(graph=G)


def test_product_never_imports_oracle():
    bad = []
    for root, _, files in os.walk(os.path.join(REPO, 'gem_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                s = open(os.path.join(root, f)).read()
                if re.search(r'^\s*(from|import)\s+.*oracle', s, flags=re.M) or 'oracle/' in s and f.endswith('.py') and 'import' in s and re.search(r'sys\.path.*oracle', s):
                    bad.append(f)
    assert not bad, bad
