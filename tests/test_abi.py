"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/mmlf_b200.h
declares, host-only helpers agree with the oracle, and the product path fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'mmlf_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(mmlf_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from mmlf_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/mmlf_b200.h but not exported'
    # and the ctypes prototype table covers the header
    assert set(names) == set(_lib.EXPORTS)
    assert _lib.lib().mmlf_abi_version() == _lib.ABI_VERSION == 5


def test_shift_taps_host_helper_matches_oracle():
    from mmlf_b200 import _lib
    l = _lib.lib()
    for disp in [2.5, -1.3, 0.0, 0.1, -0.1, 7.25, 3.0999999999999943] + list(np.arange(-3.5, 3.5, 0.1)):
        w0 = (ctypes.c_float * 9)()
        w1 = (ctypes.c_float * 9)()
        s0 = (ctypes.c_int * 9)()
        s1 = (ctypes.c_int * 9)()
        assert l.mmlf_shift_taps(float(disp), 9, w0, w1, s0, s1) == 0
        for i in range(9):
            a, b, c, d = oracle.shift_taps(float(disp), i - 4)
            assert (np.float32(w0[i]), np.float32(w1[i]), s0[i], s1[i]) == (a, b, c, d), (disp, i)


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_product_path_fails_loudly_without_gpu():
    import _fixtures as fx
    from mmlf_b200.model.feed_forward import FeedForward
    from mmlf_b200 import ops
    m = FeedForward(**fx.model_kwargs('base', False, chs=8))
    x = torch.zeros(1, 9, 3, 8, 8)
    with pytest.raises(RuntimeError):
        m(x, x, x, x)
    with pytest.raises(RuntimeError):
        ops.lf_shift(x, x, x, x, 1.0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'mmlf_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f


def test_header_is_plain_c_and_struct_layouts_match_the_ctypes_mirrors(tmp_path):
    """include/mmlf_b200.h compiles as C99 (no C++ / CUDA / torch types at the boundary) and the structs that cross it
    have the size and field offsets of their ctypes mirrors."""
    import shutil
    import subprocess
    from mmlf_b200 import _lib
    from mmlf_b200.data.augment import AugSample
    from mmlf_b200.engine import PackJob
    if shutil.which('gcc') is None:
        pytest.skip('gcc not available')
    mirrors = {'mmlf_conv_args': _lib.ConvArgs, 'mmlf_pack_job': PackJob, 'mmlf_aug_sample': AugSample}
    rename = {'in_': 'in'}
    lines = ['#include "mmlf_b200.h"', '#include <stdio.h>', '#include <stddef.h>', 'int main(void) {']
    for cname, cls in mirrors.items():
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {rename.get(fname, fname)}));')
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path / 'abi.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'abi'
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    for line in out:
        parts = line.split()
        cls = mirrors[parts[0]]
        got = [int(x) for x in parts[1:]]
        want = [ctypes.sizeof(cls)] + [getattr(cls, f).offset for f, _ in cls._fields_]
        assert got == want, f'{parts[0]}: C layout {got} != ctypes mirror {want}'
