"""The C-ABI library loads and exports every symbol include/jrystal_b200.h declares (no
compute calls: this runs without a GPU), and the ctypes prototypes cover the same set."""
import os
import re

from jrystal_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
  text = open(os.path.join(ROOT, 'include', 'jrystal_b200.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(jrb_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported():
  names = _declared()
  assert len(names) >= 20
  lib = _lib.load()
  for n in names:
    assert hasattr(lib, n), f'{n} declared in the header but not exported'
  assert sorted(_lib.SYMBOLS) == names


def test_no_gpu_calls_fail_cleanly():
  """Pure queries work without a device; error text is retrievable."""
  lib = _lib.load()
  assert lib.jrb_version() >= 100
  assert lib.jrb_launch_count() >= 0
  assert lib.jrb_plan_num_g(None) == -1
  assert lib.jrb_plan_destroy(None) == 0


def test_plan_desc_layout_matches_header():
  import ctypes
  # int32 x6, 3 pointers, int32 x2 (with natural alignment)
  assert ctypes.sizeof(_lib.PlanDesc) == 6 * 4 + 3 * 8 + 2 * 4


def test_documents_name_only_declared_entry_points():
  """INTEGRATION.md / DESIGN.md / README.md may only name C entry points the header declares
  (planned ones are listed here explicitly)."""
  planned = {'jrb_comm_create',   # peer all-reduce set-up, on a branch
             'jrb_xla_ffi'}       # name of the XLA-FFI shim library in INTEGRATION.md
  declared = set(_declared())
  text = open(os.path.join(ROOT, 'include', 'jrystal_b200.h')).read()
  types = set(re.findall(r'\b(jrb_[a-z0-9_]+)\b', text)) - declared   # typedefs / structs / enums
  for doc in ('INTEGRATION.md', 'DESIGN.md', 'README.md'):
    body = open(os.path.join(ROOT, doc)).read()
    for name in set(re.findall(r'\b(jrb_[a-z0-9_]+)\b', body)):
      base = name.rstrip('_')
      assert (name in declared or name in types or name in planned
              or any(d.startswith(base) for d in declared)), f'{doc} names {name}, not in the header'
