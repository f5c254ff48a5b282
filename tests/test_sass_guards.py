"""CPU checks on the compiled sm_100a code (cuobjdump works without a GPU): the tensor-core kernels really are
tcgen05 + TMA code, the shared-memory kernels use cp.async, and the bandwidth kernels keep their working set in registers
(a run-time index into a register array silently moves it to local memory: lf_shift lost 10 % of the HBM rate that way)."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'mmlf_b200', 'libmmlf_b200.so')


@pytest.fixture(scope='module')
def sass():
    if shutil.which('cuobjdump') is None or not os.path.exists(LIB):
        pytest.skip('cuobjdump or the built library is not available')
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    ops = collections.defaultdict(collections.Counter)
    name = None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m and name:
            ops[name][m.group(1).split('.')[0]] += 1
    assert ops, 'no sm_100a code found in the library'
    return ops


def _kernels(sass, fragment):
    hit = {k: v for k, v in sass.items() if fragment in k}
    assert hit, f'no kernel matching {fragment}'
    return hit


def test_conv_kernels_are_tcgen05_and_tma(sass):
    for frag in ('conv2x2_tc2_kernel', 'conv2x2_wgrad2_kernel'):
        for name, c in _kernels(sass, frag).items():
            assert c['UTCHMMA'] > 0, f'{name}: no tcgen05.mma (UTCHMMA)'
            assert c['UTMALDG'] > 0, f'{name}: no TMA loads (UTMALDG)'
            assert c['UTCBAR'] > 0, f'{name}: no tcgen05.commit (UTCBAR)'
            assert c['LDTM'] > 0, f'{name}: no TMEM loads (LDTM)'
            assert c['HMMA'] == 0, f'{name}: legacy mma.sync code'


def test_shared_memory_kernels_use_cp_async(sass):
    for frag in ('dpp_head_smem_kernel', 'loss_ce_smem_kernel'):
        for name, c in _kernels(sass, frag).items():
            assert c['LDGSTS'] > 0, f'{name}: no cp.async (LDGSTS)'


@pytest.mark.parametrize('frag', ['lf_shift_vec_kernel', 'slot_map_kernel', 'col_reduce_kernel', 'pack_views_vec_kernel',
                                  'lf_extract_kernel', 'dpp_head_smem_kernel', 'loss_ce_smem_kernel', 'adam_kernel',
                                  'pack_conv_weights_batch_kernel', 'wgrad_reduce_canon_kernel'])
def test_no_local_memory_in_hot_kernels(sass, frag):
    for name, c in _kernels(sass, frag).items():
        assert c['LDL'] == 0 and c['STL'] == 0, f'{name}: {c["LDL"]} LDL / {c["STL"]} STL (register array in local memory)'


def test_conv_epilogue_variants_do_not_spill(sass):
    """The specialised epilogue instances of the conv kernel (template argument = feature mask) keep everything in
    registers: at most the two register saves around the out-of-line mbarrier slow path.  The catch-all instance
    (kFGeneric = 512) and the opt-in fused BatchNorm-backward instance (2080) are allowed to spill."""
    seen = 0
    for name, c in _kernels(sass, 'conv2x2_tc2_kernel').items():
        mask = int(re.search(r'ILj(\d+)E', name).group(1))
        if mask in (512, 2080):
            continue
        seen += 1
        assert c['LDL'] <= 2 and c['STL'] <= 2, f'{name}: {c["LDL"]} LDL / {c["STL"]} STL'
    assert seen >= 10
