"""`python -m jrystal_b200 -m energy|band -c config.yaml`: the reference's command line
(`jrystal -m energy|band -c config.yaml`, jrystal/main.py) on the B200 drivers."""
import argparse
import os
import sys

import numpy as np

from . import calc
from .calc import ground_state_io
from .config import get_config


def main(argv=None):
  ap = argparse.ArgumentParser(prog='jrystal_b200')
  ap.add_argument('-m', '--mode', choices=['energy', 'band'], default='energy')
  ap.add_argument('-c', '--config', default=None,
                  help='config file with the reference keys; default: ./config.yaml if it exists '
                       '(the reference\'s default, main.py:24-29), else the built-in defaults')
  ap.add_argument('-l', '--load', default=None,
                  help='band mode: ground_state.npz (or its directory) written by the energy mode '
                       'into save_dir, instead of minimising the energy again (main.py:31-38)')
  args = ap.parse_args(argv)
  if args.config is None and os.path.exists('config.yaml'):
    args.config = 'config.yaml'
  config = get_config(args.config)
  log = print if config.verbose else None
  if args.mode == 'energy':
    out = calc.energy(config, log=log)
    print(f'Hartree Energy: {out.energies["hartree"]:.4f} Ha')
    if 'external_local' in out.energies:
      print(f'External (local) Energy: {out.energies["external_local"]:.4f} Ha')
      print(f'External (nonlocal) Energy: {out.energies["external_nonlocal"]:.4f} Ha')
    else:
      print(f'External Energy: {out.energies["external"]:.4f} Ha')
    print(f'XC Energy: {out.energies["xc"]:.4f} Ha')
    print(f'Kinetic Energy: {out.energies["kinetic"]:.4f} Ha')
    print(f'Nuclear repulsion Energy: {out.energies["ewald"]:.4f} Ha')
    print(f'Total Energy: {out.total_energy:.4f} Ha')
    print(f'{"Converged" if out.converged else "Did not converge"} after {out.steps} steps, '
          f'{out.seconds_per_step * 1e3:.3f} ms/step')
    if config.save_dir:
      os.makedirs(config.save_dir, exist_ok=True)  # before the first write: a converged run is not lost
      np.save(os.path.join(config.save_dir, 'density.npy'), out.density.cpu().numpy())
      print(f'saved {ground_state_io.save(out, config.save_dir)}')
  else:
    ground_state = ground_state_io.load(args.load, config) if args.load else None
    out = calc.band(config, ground_state=ground_state, log=log)
    name = ''.join(out.ground_state.crystal.symbols) + '_band_structure.npy'
    np.save(name, out.eigenvalues)
    print(f'saved {name}: eigenvalues {out.eigenvalues.shape} (spin, k, band)')
  return 0


if __name__ == '__main__':
  sys.exit(main())
