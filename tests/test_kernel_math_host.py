"""The arithmetic of the device FFT headers, compiled for the HOST with g++ and run on the CPU
(`-m "not gpu"` coverage of kernel code): jrystal_b200/csrc/dft_small.cuh (in-register DFTs of
radix 2..16 built from radix-2/3/4/5/7 butterflies) and fft_lines.cuh (two-stage Stockham line
FFT; every planned length 2..256, both directions, and the inverse -> forward register chaining of
the H-apply x pass) against a long-double naive DFT.  The sources are tests/host/*.cpp; they
emulate the thread loop of a CTA sequentially per phase, so what runs is the device code path
itself, not a restatement."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_INC = '/usr/local/cuda/include'


def _build_and_run(name, tmp_path):
  gxx = shutil.which('g++')
  if gxx is None or not os.path.isdir(CUDA_INC):
    pytest.skip('g++ or the CUDA headers are not available')
  exe = str(tmp_path / name)
  src = os.path.join(HERE, 'host', name + '.cpp')
  subprocess.run([gxx, '-O1', '-std=c++17', '-I', CUDA_INC, src, '-o', exe], check=True,
                 timeout=300)
  run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
  assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-2000:]
  return run.stdout.strip().splitlines()


@pytest.mark.parametrize('name', ['test_dft_small', 'test_fft_lines'])
def test_device_fft_headers_on_the_host(name, tmp_path):
  lines = _build_and_run(name, tmp_path)
  assert lines[-1] == 'OK'
  assert not any('BAD' in ln or 'FAIL' in ln for ln in lines)
  if name == 'test_fft_lines':
    covered = {int(ln.split('=')[1].split('(')[0]) for ln in lines if ln.startswith('N=')}
    # every line length the pencil passes are compiled for
    for n in (7, 8, 9, 12, 16, 24, 32, 40, 45, 48, 49, 50, 54, 56, 60, 64, 72, 80, 81, 90, 96, 100,
              112, 128):
      assert n in covered, n


def test_device_xc_functionals_on_the_host(tmp_path):
  """jrystal_b200/csrc/xc_functionals.cuh (lda_x, lda_x+lda_c_pw, gga_x_pbe, gga_x_pbe+gga_c_pbe:
  eps_xc and the derivatives the fused grid kernels use, the GGA ones by forward-mode duals)
  compiled for the host, against the oracle's functionals differentiated by torch autograd, from
  the density thresholds up to 900 electrons / bohr^3."""
  import torch
  from oracle import reference_port as rp
  lines = _build_and_run('test_xc_functionals', tmp_path)
  lda = {1: ['lda_x'], 2: ['lda_x', 'lda_c_pw']}
  gga = {3: ['gga_x_pbe'], 4: ['gga_x_pbe', 'gga_c_pbe']}
  fns = {'lda_x': rp._eps_lda_x, 'lda_c_pw': rp._eps_lda_c_pw,
         'gga_x_pbe': rp._eps_gga_x_pbe, 'gga_c_pbe': rp._eps_gga_c_pbe}
  n_lda = n_gga = n_pol = 0
  for ln in lines:
    tok = ln.split()
    vals = [float(t) for t in tok[2:]]
    xc_id = int(tok[1])
    if tok[0] == 'pol':
      # two spin channels (xc.py:54-64): spin-scaled exchange + the polarised PW92 correlation
      ru, rd, eps, d_up, d_dn = vals
      rho = torch.tensor([[ru], [rd]], dtype=torch.float64, requires_grad=True)
      parts = [rp._lda(f, rho) for f in lda[xc_id]]
      e = sum(parts)
      (g,) = torch.autograd.grad(e.sum(), rho, retain_graph=True)
      scale = sum(torch.autograd.grad(p_.sum(), rho, retain_graph=True)[0].abs().max().item()
                  for p_ in parts)
      assert abs(eps - e.item()) <= 1e-12 * abs(e.item()), ln
      # an empty channel: the exchange derivative of that channel is switched off below the
      # threshold on both sides; (1 -+ zeta)^(1/3) has an infinite slope there, not compared
      if ru > 1e-15:
        assert abs(d_up - g[0].item()) <= 1e-9 * scale, ln
      if rd > 1e-15:
        assert abs(d_dn - g[1].item()) <= 1e-9 * scale, ln
      n_pol += 1
      continue
    if tok[0] == 'lda':
      n, eps, deps = vals
      x = torch.tensor([n], dtype=torch.float64, requires_grad=True)
      e = sum(fns[f](x) for f in lda[xc_id])
      (de,) = torch.autograd.grad(e.sum(), x)
      # below 1e-6 the oracle's log(1 + 1/Q) of PW92 loses digits the kernel's log1p keeps
      tol = 1e-13 if n >= 1e-6 else 1e-9
      assert abs(eps - e.item()) <= tol * max(abs(e.item()), 1e-300), ln
      if n > 1e-15:   # at and below the threshold both sides return eps = 0 and no derivative
        assert abs(deps - de.item()) <= (1e-11 if n >= 1e-6 else 1e-8) * abs(de.item()), ln
      else:
        assert eps == 0.0 and deps == 0.0
      n_lda += 1
    else:
      rho, sigma, eps, d_rho, d_sigma = vals
      x = torch.tensor([rho], dtype=torch.float64, requires_grad=True)
      sg = torch.tensor([sigma], dtype=torch.float64, requires_grad=True)
      parts = [fns[f](x, sg) for f in gga[xc_id]]
      e = sum(parts)
      dr, ds = torch.autograd.grad(e.sum(), [x, sg], retain_graph=True)
      # exchange and correlation derivatives cancel to ~1e-14 of their size at very low density:
      # the scale of a derivative is the sum of the magnitudes of its parts
      each = [torch.autograd.grad(p_.sum(), [x, sg], retain_graph=True) for p_ in parts]
      scale_r = sum(abs(g_[0].item()) for g_ in each)
      scale_s = sum(abs(g_[1].item()) for g_ in each)
      tol = 1e-12 if rho >= 1e-6 else 1e-9
      assert abs(eps - e.item()) <= tol * max(abs(e.item()), 1e-300), ln
      if rho > 1e-15:  # (gga_c_pbe switches on above ITS threshold 1e-12 on both sides)
        assert abs(d_rho - dr.item()) <= 1e2 * tol * max(scale_r, 1e-300), ln
        assert abs(d_sigma - ds.item()) <= 1e2 * tol * max(scale_s, 1e-300), ln
      n_gga += 1
  assert n_lda == 24 and n_gga == 132 and n_pol == 84
