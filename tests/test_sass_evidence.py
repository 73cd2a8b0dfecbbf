"""The built library really contains what DESIGN.md says it does: tcgen05 MMAs with TMEM loads in the two dense kernels (the
coarsest quadtree level is ON the benched path), TMA tensor loads + mbarriers in the tile kernels, packed FFMA2 in the gather
kernels.  Runs cuobjdump on the in-tree .so (no GPU needed)."""
import re
import shutil
import subprocess

import pytest

from casmtr_b200 import _lib


@pytest.fixture(scope='module')
def sass():
    exe = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    try:
        r = subprocess.run([exe, '-sass', _lib.LIB_PATH], capture_output=True, text=True, timeout=300)
    except (OSError, subprocess.TimeoutExpired):
        pytest.skip('cuobjdump not available')
    if r.returncode != 0 or 'Function :' not in r.stdout:
        pytest.skip('cuobjdump could not read the library')
    per = {}
    cur = None
    for line in r.stdout.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            per.setdefault(cur, [])
        elif cur is not None:
            per[cur].append(line)
    return {k: '\n'.join(v) for k, v in per.items()}


def _kernels(sass, name):
    out = [body for fn, body in sass.items() if name in fn]
    assert out, f'no kernel named *{name}* in the library'
    return out


def test_library_is_sm100a_only(sass):
    r = subprocess.run([shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump', '-lelf', _lib.LIB_PATH], capture_output=True, text=True)
    archs = set(re.findall(r'sm_(\d+a?)', r.stdout))
    assert archs == {'100a'}, archs


@pytest.mark.parametrize('kernel', ['qtatt_coarse_tc_kernel', 'coarse_rowstats_kernel'])
def test_dense_kernels_run_on_tcgen05(sass, kernel):
    for body in _kernels(sass, kernel):
        assert 'UTCHMMA' in body, f'{kernel}: no tcgen05.mma'
        assert 'LDTM' in body, f'{kernel}: no tcgen05.ld (TMEM -> registers)'
        assert 'UTMALDG' in body, f'{kernel}: operands are not TMA-fed'
        assert 'SYNCS' in body, f'{kernel}: no mbarrier'


@pytest.mark.parametrize('kernel', ['cascade_att_tile_kernel', 'cascade_match_tile_kernel'])
def test_tile_kernels_use_tma_and_mbarriers(sass, kernel):
    for body in _kernels(sass, kernel):
        assert 'UTMALDG' in body and 'SYNCS' in body and 'FFMA2' in body


@pytest.mark.parametrize('kernel', ['quad_cta_kernel', 'quad_attention_kernel'])
def test_gather_kernels_use_async_copies_and_packed_fma(sass, kernel):
    for body in _kernels(sass, kernel):
        assert 'LDGSTS' in body and 'FFMA2' in body
        assert 'UTCHMMA' not in body            # M = 4: no tensor cores here, by design
