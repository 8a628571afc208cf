"""Import the UNMODIFIED reference module /root/reference/dreamer4/dreamer4.py with the
third-party shims in oracle/shims on sys.path.  Build-container only (the GPU box has no
/root/reference); used by oracle/make_golden.py and by tests that are skipped when the
reference tree is absent.  Test infrastructure."""
import importlib.util
import os
import sys

REF_ROOT = os.environ.get('D4_REFERENCE_ROOT', '/root/reference')
SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shims')

def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, 'dreamer4', 'dreamer4.py'))

def import_reference():
    """Returns the reference's `dreamer4.dreamer4` module object (not the package: the
    package __init__ pulls in trainers -> accelerate, which is not installed)."""
    name = '_d4_reference_module'
    if name in sys.modules:
        return sys.modules[name]
    assert reference_available(), f'no reference tree at {REF_ROOT}'
    if SHIMS not in sys.path:
        sys.path.insert(0, SHIMS)
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, 'dreamer4', 'dreamer4.py'))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
