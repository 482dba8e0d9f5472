#!/usr/bin/env python
"""Generates the golden vectors tests/golden/*.npz from the ORACLE (oracle/reference_port.py).

The reference (sail-sg/jrystal) cannot be imported in this environment (jax, jax_xc, ase are
absent and not installable), and it ships no literal energy / density / gradient vectors
(SURVEY.md 8c), so these fixtures are outputs of the CPU restatement, which tests/test_oracle.py
pins against the reference's own test identities and crystal goldens first.  Each file holds
the seeded inputs' recipe (structure, grid, k-grid, mask, seed) and the outputs
(E_kin, E_ext, E_har, E_xc, rho, dE/dw_re, dE/dw_im, dE/docc, band trace data).

  python tests/golden/make_golden.py [case ...]   # rewrites the .npz files (deterministic)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import reference_port as rp  # noqa: E402

# name -> recipe; the first is the reference's own unit-test fixture (pw_test.py:25-34)
CASES = {
  'diamond_789_cubic': dict(name='diamond', grid=[7, 8, 9], kgrid=[2, 2, 1], mask='cubic',
                            cutoff=None, nb=12, seed=123, xc='lda_x'),
  'diamond_16_sph': dict(name='diamond', grid=[16, 16, 16], kgrid=[1, 1, 2], mask='spherical',
                         cutoff=20.0, nb=10, seed=7, xc='lda_x'),
  'si_24x32x48_pw': dict(name='si', grid=[24, 32, 48], kgrid=[1, 1, 1], mask='spherical',
                         cutoff=8.0, nb=9, seed=11, xc='lda_x+lda_c_pw'),
  # the GGA branch (xc.py:67-112), the reference's config.yaml default functional
  'diamond_12_pbe': dict(name='diamond', grid=[12, 12, 12], kgrid=[1, 1, 2], mask='spherical',
                         cutoff=10.0, nb=6, seed=5, xc='gga_x_pbe+gga_c_pbe'),
}


def inputs(c):
  s = rp.System.from_name(c['name'], c['grid'], c['kgrid'], c['cutoff'], mask_method=c['mask'])
  p = rp.param_init(c['seed'], c['nb'], s.num_k, s.mask)
  occ = rp.occupation_uniform(s.num_k, s.num_electrons, num_bands=c['nb']).numpy()
  occ = occ * (1.0 + 0.1 * np.random.default_rng(c['seed'] + 1).random(occ.shape))
  return s, p['w_re'], p['w_im'], occ


def main():
  only = sys.argv[1:]
  for key, c in CASES.items():
    if only and key not in only:
      continue
    s, w_re, w_im, occ = inputs(c)
    ref = rp.energy_and_grad(s, w_re, w_im, occ, xc=c['xc'], occ_grad=True)
    band = rp.band_trace_and_grad(s, w_re, w_im, ref['density'], xc=c['xc'])
    out = dict(
      energies=np.array([ref['e_kin'], ref['e_ext'], ref['e_har'], ref['e_xc']]),
      density=ref['density'], g_re=ref['g_re'], g_im=ref['g_im'], g_occ=ref['g_occ'],
      band_per_band=band['per_band'], band_g_re=band['g_re'], band_g_im=band['g_im'],
      # input digests: the recipe must regenerate exactly these inputs
      w_re_sum=np.array(w_re.sum()), w_im_sum=np.array(w_im.sum()), occ=occ,
      num_g=np.array(s.num_g), vol=np.array(s.vol),
    )
    path = os.path.join(HERE, key + '.npz')
    np.savez_compressed(path, **out)
    print(key, 'ng', s.num_g, 'E', out['energies'].sum(), os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()
