"""CPU: the C-ABI library builds, loads and exports every symbol include/selfocc_b200.h declares."""
import os
import re
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'selfocc_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(so_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from selfocc_b200 import build, _lib
    build.build()
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 12
    nm = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(l.split()[-1] for l in nm.splitlines() if l.strip())
    for s in declared:
        assert s in exported, 'header declares %s but the library does not export it' % s
        assert s in _lib.SIGNATURES, 'ctypes binding lacks %s' % s
    assert sorted(_lib.SIGNATURES) == declared
    assert lib.so_abi_version() == _lib.ABI_VERSION
    assert lib.so_error_string(-1) == b'invalid argument'


def test_sm100a_only():
    from selfocc_b200 import _lib, build
    build.build()
    out = subprocess.run(['cuobjdump', '--list-elf', _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_\d+a?', out))
    assert archs == {'sm_100a'}, archs


def test_ops_refuse_cpu_tensors():
    import torch
    from selfocc_b200 import ops
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.msda_forward(torch.zeros(1, 4, 1, 16), torch.tensor([[2, 2]]), torch.tensor([0]),
                         torch.zeros(1, 1, 1, 1, 1, 2), torch.zeros(1, 1, 1, 1, 1))
