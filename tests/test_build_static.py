"""Static properties of the built library that the measured performance depends on, read with
cuobjdump (no GPU): the hot kernels of the BASELINE configurations keep their register / spill
budget (two or more resident CTAs, nothing in local memory on the C2 path), and the dense
contractions are on the FP64 tensor instruction.  Guards the numbers of profiles/r01_kernel_tables.md
against a silent change of the build (flags, launch bounds, a refactored header)."""
import os
import re
import shutil
import subprocess

import pytest

from jrystal_b200 import _lib

pytestmark = pytest.mark.skipif(shutil.which('cuobjdump') is None, reason='no cuobjdump')


def _resources():
  out = subprocess.run(['cuobjdump', '-res-usage', _lib.LIB_PATH], capture_output=True, text=True,
                       check=True).stdout
  rows = re.findall(r'Function (\S+):\n\s+REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)', out)
  names = subprocess.run(['c++filt'], input='\n'.join(r[0] for r in rows), capture_output=True,
                         text=True, check=True).stdout.splitlines()
  res = {}
  for n, r in zip(names, rows):
    n = re.sub(r'^void ', '', n).replace('jrb::', '')
    res[n.split('(')[0]] = dict(reg=int(r[1]), stack=int(r[2]), smem=int(r[3]), local=int(r[4]))
  return res


@pytest.fixture(scope='module')
def res():
  return _resources()


def test_library_is_sm_100a_only():
  out = subprocess.run(['cuobjdump', '-lelf', _lib.LIB_PATH], capture_output=True, text=True,
                       check=True).stdout
  archs = set(re.findall(r'sm_(\d+a?)', out))
  assert archs == {'100a'}, archs


def test_c2_hot_kernels_do_not_spill(res):
  """C2 / C4 (64 x 64 planes, 49-point z lines): the fused plane kernels and the z passes hold
  everything in registers; 128 registers x 256 threads leaves two CTAs per SM."""
  for name in ('k_yx_vmul<64, true, true>', 'k_yx_density<64, true, true>',
               'k_z_inv_scatter<49>', 'k_z_fwd_gather<49>'):
    r = res[name]
    assert r['stack'] == 0 and r['local'] == 0, (name, r)
    assert r['reg'] <= 128, (name, r)


def test_c3_hot_kernels_stay_within_their_budget(res):
  """C3 (81 x 81 and 128 x 128 planes): launch bounds cap the kernels at 128 registers; the price
  is a few spilled values (measured trade: profiles/r01_optimization_log.md), never more than 64 B."""
  for name in ('k_yx_vmul<81, true, false>', 'k_yx_density<81, true, false>', 'k_yx128_vmul',
               'k_yx128_density', 'k_z_inv_scatter<81>', 'k_z_fwd_gather<81>',
               'k_z_inv_scatter<128>', 'k_z_fwd_gather<128>'):
    r = res[name]
    assert r['reg'] <= 128 and r['stack'] <= 64 and r['local'] == 0, (name, r)


def test_qr_products_hold_their_tiles_in_registers(res):
  for name, r in res.items():
    if name.startswith('k_gram<') or name.startswith('k_apply<'):
      assert r['stack'] == 0 and r['local'] == 0, (name, r)
  # nb <= 32 variants: three resident CTAs (commit 0dff668)
  assert res['k_gram<6, 2, 0>']['reg'] <= 80
  assert res['k_apply<0, 4, 2, false>']['reg'] <= 72


def test_dense_contractions_use_the_fp64_tensor_instruction():
  """DMMA in the Gram / apply kernels, asynchronous copies (LDGSTS) in their operand rings."""
  out = subprocess.run(['cuobjdump', '-sass', '-fun', 'k_gram', _lib.LIB_PATH],
                       capture_output=True, text=True)
  if out.returncode != 0 or 'DMMA' not in out.stdout:
    # -fun wants the mangled name on some versions: fall back to one object of the build tree
    obj = os.path.join(os.path.dirname(_lib.LIB_PATH), 'build', 'qr.o')
    if not os.path.exists(obj):
      pytest.skip('no per-object build tree')
    out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True, check=True)
  assert out.stdout.count('DMMA') > 100
  assert 'LDGSTS' in out.stdout


def test_fused_plane_kernels_stage_with_tma():
  """The fused y+x plane kernels stage their band-planes with TMA tensor copies completed on an
  mbarrier (north_star (1): transposes staged through TMA / shared memory): the shipped SASS holds
  UTMALDG (cp.async.bulk.tensor) and SYNCS (mbarrier) instructions, and the cp.async fallback."""
  objs = [os.path.join(os.path.dirname(_lib.LIB_PATH), 'build', f'fft_fused_g{i}.o') for i in range(3)]
  if not all(os.path.exists(o) for o in objs):
    pytest.skip('no per-object build tree')
  tma = mbar = ldgsts = 0
  for obj in objs:
    out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True, check=True).stdout
    tma += out.count('UTMALDG')
    mbar += out.count('SYNCS')
    ldgsts += out.count('LDGSTS')
  assert tma > 0 and mbar > 0 and ldgsts > 0, (tma, mbar, ldgsts)
