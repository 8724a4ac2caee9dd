"""The C-ABI shared library: it builds for sm_100a without a GPU, loads, exports every symbol include/vit_b200.h
declares, and answers the GPU-free entry points (sizes, algorithm selection, argument validation, error strings)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'vit_b200.h')


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(vit_[a-z0-9_]+)\s*\(', text)))


def test_header_and_binding_list_agree(cuda_lib):
    from viterbi_spl_b200 import _lib
    assert declared_functions() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol(cuda_lib):
    for name in declared_functions():
        assert hasattr(cuda_lib, name), f'{name} is declared in vit_b200.h but not exported'


def test_sm100a_code_is_embedded():
    from viterbi_spl_b200 import _lib
    import subprocess
    out = subprocess.run(['cuobjdump', '-lelf', _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip('cuobjdump not available')
    assert 'sm_100a' in out.stdout


def test_version_and_strerror(cuda_lib):
    assert cuda_lib.vit_version() >= 100
    seen = set()
    for code in range(0, -8, -1):
        msg = cuda_lib.vit_strerror(code).decode()
        assert msg
        seen.add(msg)
    assert len(seen) >= 7
    assert cuda_lib.vit_last_cuda_error().decode() == ''
    assert cuda_lib.vit_launch_count() == 0 or cuda_lib.vit_launch_count() > 0


def test_workspace_and_algo_selection(cuda_lib):
    from viterbi_spl_b200 import _lib
    # every pitch-bin state set (321 / 361 / 722) fits the tensor-memory kernel; beyond 8 x 192 states the generic
    # backpointer kernel takes over
    assert cuda_lib.vit_select_algo(1024, 3000, 361) == _lib.ALGO_TMEM
    assert cuda_lib.vit_select_algo(1, 3000, 321) == _lib.ALGO_TMEM
    assert cuda_lib.vit_select_algo(16, 100, 722) == _lib.ALGO_TMEM
    assert cuda_lib.vit_select_algo(16, 100, 1537) == _lib.ALGO_BACKPOINTER
    n = _lib.workspace_bytes(1024, 3000, 361, _lib.ALGO_TMEM)
    assert 1024 * 3000 * 361 * 4 <= n <= 1024 * 3000 * 361 * 4 + (1 << 20)       # fp32 delta history + packed logA^T
    n = _lib.workspace_bytes(1024, 3000, 361, _lib.ALGO_CLUSTER)
    assert 1024 * 3000 * 361 * 4 <= n <= 1024 * 3000 * 361 * 4 + (1 << 20)       # fp32 delta history + packed logA^T
    n = _lib.workspace_bytes(1024, 3000, 361, _lib.ALGO_BACKPOINTER)
    assert 1024 * 3000 * 361 * 2 <= n <= 1024 * 3000 * 361 * 2 + (1 << 20)       # uint16 backpointers
    out = ctypes.c_size_t(0)
    assert cuda_lib.vit_workspace_bytes(4, 10, 722, _lib.ALGO_CLUSTER, ctypes.byref(out)) == -4
    assert cuda_lib.vit_workspace_bytes(4, 0, 10, 0, ctypes.byref(out)) == -1
    assert cuda_lib.vit_workspace_bytes(4, 10, 70000, 0, ctypes.byref(out)) == -2
    assert cuda_lib.vit_workspace_bytes(4, 10, 10, 0, None) == -1


def test_decode_rejects_bad_arguments_without_touching_the_gpu(cuda_lib):
    one = ctypes.c_void_p(256)      # any aligned non-null value: rejected before it is dereferenced
    assert cuda_lib.vit_decode_f32(None, one, one, None, 1, 1, 1, one, 1024, one, None, None) == -1
    assert cuda_lib.vit_decode_f32(one, one, one, None, -1, 1, 1, one, 1024, one, None, None) == -1
    assert cuda_lib.vit_decode_f32(one, one, one, None, 1, 0, 1, one, 1024, one, None, None) == -1
    assert cuda_lib.vit_decode_f32(one, one, one, None, 1, 1, 65536, one, 1024, one, None, None) == -2
    assert cuda_lib.vit_decode_f32(one, one, one, None, 1, 1, 8, ctypes.c_void_p(257), 1 << 20, one, None, None) == -6
    assert cuda_lib.vit_decode_f32(one, one, one, None, 4, 100, 361, one, 16, one, None, None) == -3


def test_forward_backward_ex_rejects_bad_arguments_without_touching_the_gpu(cuda_lib):
    from viterbi_spl_b200 import _lib
    one = ctypes.c_void_p(256)
    opts = _lib.FbOpts()
    opts.impl = _lib.FB_BANDED
    st = _lib.Structure(1, 14, 360, 0.0, 0.0)
    opts.structure = ctypes.pointer(st)
    f = cuda_lib.vit_forward_backward_f32_ex
    assert f(None, one, one, None, 1, 1, 361, one, 1 << 30, one, None, ctypes.byref(opts), None) == -1
    assert f(one, one, one, None, 1, 0, 361, one, 1 << 30, one, None, ctypes.byref(opts), None) == -1
    assert f(one, one, one, None, 1, 1, 70000, one, 1 << 30, one, None, ctypes.byref(opts), None) == -2
    assert f(one, one, one, None, 4, 100, 361, ctypes.c_void_p(257), 1 << 30, one, None, ctypes.byref(opts), None) == -6
    # the banded kernels need the structure: none / a log-domain background / a band wider than any instance -> unsupported
    opts.structure = None
    assert f(one, one, one, None, 4, 100, 361, one, 1 << 30, one, None, ctypes.byref(opts), None) == -4
    for bad in (_lib.Structure(1, 14, 360, -87.3, 0.0), _lib.Structure(0, 14, 360, 0.0, 0.0), _lib.Structure(1, 57, 721, 0.0, 0.0)):
        opts.structure = ctypes.pointer(bad)
        S = 722 if bad.halfwidth > 14 else 361
        assert f(one, one, one, None, 4, 100, S, one, 1 << 30, one, None, ctypes.byref(opts), None) == -4
    # the workspace covers the dense kernels' needs plus the structured kernels' normalisers and parameter block
    n = ctypes.c_size_t(0)
    assert cuda_lib.vit_fb_workspace_bytes(64, 100, 361, ctypes.byref(n)) == 0
    assert n.value >= 64 * 100 * 4 * 2
