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


@pytest.mark.parametrize('name', ['test_dft_small', 'test_fft_lines'])
def test_device_fft_headers_on_the_host(name, tmp_path):
  gxx = shutil.which('g++')
  if gxx is None or not os.path.isdir(CUDA_INC):
    pytest.skip('g++ or the CUDA headers are not available')
  exe = str(tmp_path / name)
  src = os.path.join(HERE, 'host', name + '.cpp')
  subprocess.run([gxx, '-O1', '-std=c++17', '-I', CUDA_INC, src, '-o', exe], check=True,
                 timeout=300)
  run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
  assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-2000:]
  lines = run.stdout.strip().splitlines()
  assert lines[-1] == 'OK'
  assert not any('BAD' in ln or 'FAIL' in ln for ln in lines)
  if name == 'test_fft_lines':
    covered = {int(ln.split('=')[1].split('(')[0]) for ln in lines if ln.startswith('N=')}
    # every line length the pencil passes are compiled for
    for n in (7, 8, 9, 12, 16, 24, 32, 40, 45, 48, 49, 50, 54, 56, 60, 64, 72, 80, 81, 90, 96, 100,
              112, 128):
      assert n in covered, n
