"""ctypes binding of libjrystal_b200.so (the C ABI of include/jrystal_b200.h).

There is no CPU fallback: importing this module without the built CUDA library raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libjrystal_b200.so')

XC_IDS = {'lda_x': 1, 'lda_x+lda_c_pw': 2, 'gga_x_pbe': 3, 'gga_x_pbe+gga_c_pbe': 4}
FFT_FORWARD, FFT_INVERSE = -1, 1
# axis lengths with compiled pencil passes; FUSED: also the fused y+x plane kernels (nx == ny)
LINE_LENGTHS = (7, 8, 9, 12, 16, 24, 32, 36, 40, 45, 48, 49, 50, 54, 56, 60, 64, 72, 80, 81, 90, 96,
                100, 112, 128)
FUSED_LENGTHS = (7, 8, 9, 12, 16, 24, 32, 48, 49, 64, 72, 81, 96, 128)


class PlanDesc(ctypes.Structure):
  _fields_ = [
    ('nx', ctypes.c_int32), ('ny', ctypes.c_int32), ('nz', ctypes.c_int32),
    ('ns', ctypes.c_int32), ('nk', ctypes.c_int32), ('nb', ctypes.c_int32),
    ('mask', ctypes.c_void_p), ('kpts', ctypes.c_void_p), ('cell', ctypes.c_void_p),
    ('device', ctypes.c_int32), ('batch_groups', ctypes.c_int32),
  ]


# name -> (restype, argtypes); every symbol include/jrystal_b200.h declares
_P, _I32, _I64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
SYMBOLS = {
  'jrb_plan_create': (ctypes.c_int, [ctypes.POINTER(PlanDesc), ctypes.POINTER(_P)]),
  'jrb_plan_destroy': (ctypes.c_int, [_P]),
  'jrb_plan_num_g': (_I64, [_P]),
  'jrb_plan_workspace_bytes': (_I64, [_P]),
  'jrb_plan_set_orbital_grid': (ctypes.c_int, [_P, _I32, _I32, _I32]),
  'jrb_plan_orbital_grid': (ctypes.c_int, [_P, _P]),
  'jrb_plan_min_orbital_grid': (ctypes.c_int, [_P, _P]),
  'jrb_plan_orbital_fused': (ctypes.c_int, [_P]),
  'jrb_plan_psi_cache_bytes': (ctypes.c_int64, [_P]),
  'jrb_plan_phase_timing': (ctypes.c_int, [_P, ctypes.c_int32]),
  'jrb_plan_phase_times': (ctypes.c_int, [_P, _P]),
  'jrb_set_atoms': (ctypes.c_int, [_P, _P, _P, _I32, _P]),
  'jrb_set_external_potential': (ctypes.c_int, [_P, _P, _P]),
  'jrb_external_position_gradient': (ctypes.c_int, [_P, _P, _P, _P]),
  'jrb_set_nonlocal': (ctypes.c_int, [_P, _P, _I32, _P]),
  'jrb_nonlocal_energy': (ctypes.c_int, [_P, _P, _P, _P, _P]),
  'jrb_set_kpoints': (ctypes.c_int, [_P, _P, _P]),
  'jrb_qr_fwd': (ctypes.c_int, [_P, _P, _P, _P, _P, _P]),
  'jrb_qr_bwd': (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P]),
  'jrb_plan_create_rows': (ctypes.c_int, [_I64, _I32, _I32, _I32, _I32, ctypes.POINTER(_P)]),
  'jrb_qr_rows_gram': (ctypes.c_int, [_P, _P, _P, _I32, _P, _P]),
  'jrb_qr_rows_apply': (ctypes.c_int, [_P, _P, _P, _I32, _P, _P, _P, _P]),
  'jrb_qr_rows_bwd_gram': (ctypes.c_int, [_P, _P, _P, _P, _P]),
  'jrb_qr_rows_bwd_apply': (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
  'jrb_check_status': (ctypes.c_int, [_P, _P]),
  'jrb_expand': (ctypes.c_int, [_P, _P, _P, _P]),
  'jrb_squeeze': (ctypes.c_int, [_P, _P, _P, _P]),
  'jrb_density': (ctypes.c_int, [_P, _P, _P, _P, _P]),
  'jrb_kinetic': (ctypes.c_int, [_P, _P, _P, _P]),
  'jrb_grid_potential': (ctypes.c_int, [_P, _P, _I32, _I32, _P, _P, _P]),
  'jrb_potential': (ctypes.c_int, [_P, _P, _I32, _I32, _I32, _P, _P]),
  'jrb_density_reciprocal': (ctypes.c_int, [_P, _P, _P, _P]),
  'jrb_wave_grid': (ctypes.c_int, [_P, _P, _P, _P]),
  'jrb_hpsi': (ctypes.c_int, [_P, _P, _P, _P, _P]),
  'jrb_hpsi_prepare': (ctypes.c_int, [_P, _P, _P]),
  'jrb_band_expect': (ctypes.c_int, [_P, _P, _P, _P, _P]),
  'jrb_hamiltonian_matrix': (ctypes.c_int, [_P, _P, _P, _P, _P]),
  'jrb_fft3d': (ctypes.c_int, [_P, _P, _P, _I32, _I64, _P]),
  'jrb_eval_begin': (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P]),
  'jrb_eval_finish': (ctypes.c_int, [_P, _P, _P, _P, _I32, _P, _P, _P, _P, _P]),
  'jrb_eval': (ctypes.c_int, [_P, _P, _P, _P, _I32, _P, _P, _P, _P, _P, _P]),
  'jrb_comm_handle_bytes': (ctypes.c_int, []),
  'jrb_comm_create': (ctypes.c_int, [_P, _I32, _I32, _I64, _P]),
  'jrb_comm_connect': (ctypes.c_int, [_P, _P]),
  'jrb_comm_world': (ctypes.c_int, [_P]),
  'jrb_allreduce': (ctypes.c_int, [_P, _P, _I64, _P]),
  'jrb_allreduce_rho': (ctypes.c_int, [_P, _P, _P, _P]),
  'jrb_energy_grad_host': (ctypes.c_int, [_P, _P, _P, _P, _I32, _P, _P, _P, _P]),
  'jrb_adam_tick': (ctypes.c_int, [_P, ctypes.c_double, ctypes.c_double, _P]),
  'jrb_adam_apply': (ctypes.c_int, [_I64, _P, _P, _P, _P, ctypes.c_double, ctypes.c_double,
                                    ctypes.c_double, ctypes.c_double, _P, _P]),
  'jrb_last_error': (ctypes.c_char_p, []),
  'jrb_version': (ctypes.c_int, []),
  'jrb_launch_count': (_I64, []),
}

_lib = None


def load():
  """Load the shared library (once) and set the prototypes.  Fails loudly."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise ImportError(
      f'{LIB_PATH} is missing: build it with jrystal_b200/csrc/build.sh or '
      '`python -c "import __graft_entry__ as g; g.build()"`. There is no CPU fallback.'
    )
  lib = ctypes.CDLL(LIB_PATH)
  for name, (res, args) in SYMBOLS.items():
    fn = getattr(lib, name)
    fn.restype = res
    fn.argtypes = args
  _lib = lib
  return lib


class JrbError(RuntimeError):
  pass


def check(rc):
  if rc != 0:
    msg = load().jrb_last_error()
    raise JrbError(f'jrystal_b200 error {rc}: {msg.decode() if msg else "?"}')
