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


def test_struct_sizes_are_stable(native_lib, tmp_path):
    """ctypes mirrors == the C structs of include/gemb200.h (compiled here with gcc), field by field in size."""
    import subprocess
    from gem_b200 import _native
    src = tmp_path / 'sz.c'
    src.write_text('#include <stdio.h>\n#include "gemb200.h"\nint main(void){printf("%zu %zu %zu\\n", '
                   'sizeof(gemb_hope_opts), sizeof(gemb_hope_stats), sizeof(gemb_n2v_stats));return 0;}\n')
    exe = tmp_path / 'sz'
    subprocess.check_call(['gcc', '-I', os.path.join(REPO, 'include'), str(src), '-o', str(exe)])
    c_sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert [ctypes.sizeof(_native.HopeOpts), ctypes.sizeof(_native.HopeStats), ctypes.sizeof(_native.N2VStats)] == c_sizes
    assert c_sizes == [72, 152, 120]         # an ABI change must be deliberate (r2: + beta_used, push_bytes)


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
