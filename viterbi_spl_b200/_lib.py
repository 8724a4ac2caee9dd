"""ctypes binding of libvit_b200.so -- the C ABI declared in include/vit_b200.h.

There is NO CPU fallback: if the CUDA library is missing or fails to load, importing the decoder raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# (VIT_B200_LIB: a development build of the same library, e.g. one compiled with -DVIT_FB_STAMPS)
LIB_PATH = os.environ.get('VIT_B200_LIB') or os.path.join(_HERE, 'libvit_b200.so')

VIT_OK = 0
ALGO_AUTO, ALGO_BACKPOINTER, ALGO_CLUSTER, ALGO_TMEM, ALGO_BANDED, ALGO_STREAM = 0, 1, 2, 3, 4, 5
ALGO_NAMES = {'auto': ALGO_AUTO, 'backpointer': ALGO_BACKPOINTER, 'cluster': ALGO_CLUSTER, 'tmem': ALGO_TMEM,
              'banded': ALGO_BANDED, 'stream': ALGO_STREAM}

# every symbol include/vit_b200.h declares (tests/test_abi.py checks the built library exports them all)
EXPORTS = ['vit_version', 'vit_strerror', 'vit_last_cuda_error', 'vit_launch_count', 'vit_select_algo',
           'vit_workspace_bytes', 'vit_decode_f32', 'vit_decode_f32_ex', 'vit_upload_frames_f32',
           'vit_fb_workspace_bytes', 'vit_forward_backward_f32', 'vit_forward_backward_f32_ex', 'vit_emissions_f32', 'vit_voiced_bins',
           'vit_analyze_structure_f32', 'vit_clips_in_flight', 'vit_melody_stats_f32']


class Structure(ctypes.Structure):
    """struct vit_structure"""
    _fields_ = [('kind', ctypes.c_int32), ('halfwidth', ctypes.c_int32), ('dense_index', ctypes.c_int32),
                ('background', ctypes.c_float), ('dense_row_max', ctypes.c_float)]


class DecodeOpts(ctypes.Structure):
    """struct vit_decode_opts"""
    _fields_ = [('algo', ctypes.c_int32), ('reserved', ctypes.c_int32),
                ('d_backpointers', ctypes.c_void_p), ('d_delta', ctypes.c_void_p),
                ('ev_forward_begin', ctypes.c_void_p), ('ev_forward_end', ctypes.c_void_p),
                ('frame_begin', ctypes.c_int32), ('frame_end', ctypes.c_int32),
                ('skip_backtrace', ctypes.c_int32), ('reserved2', ctypes.c_int32),
                ('structure', ctypes.POINTER(Structure)), ('backtrace_stream', ctypes.c_void_p)]


class FbOpts(ctypes.Structure):
    """struct vit_fb_opts"""
    _fields_ = [('impl', ctypes.c_int32), ('reserved', ctypes.c_int32), ('structure', ctypes.POINTER(Structure))]


FB_AUTO, FB_TC, FB_SIMT, FB_BANDED = 0, 1, 2, 3
FB_IMPLS = {'auto': FB_AUTO, 'tc': FB_TC, 'simt': FB_SIMT, 'banded': FB_BANDED}


class VitError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f'libvit_b200: {msg} (status {code})')
        self.code = code


_lib = None


def load():
    """dlopen the library (building is the job of viterbi_spl_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f'{LIB_PATH} not found: build it with `python -m viterbi_spl_b200.build` '
                          '(needs nvcc; there is no CPU fallback)')
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    L.vit_version.restype = ci
    L.vit_strerror.restype = ctypes.c_char_p
    L.vit_strerror.argtypes = [ci]
    L.vit_last_cuda_error.restype = ctypes.c_char_p
    L.vit_launch_count.restype = ctypes.c_uint64
    L.vit_select_algo.restype = ci
    L.vit_select_algo.argtypes = [ci, ci, ci]
    L.vit_workspace_bytes.restype = ci
    L.vit_workspace_bytes.argtypes = [ci, ci, ci, ci, ctypes.POINTER(sz)]
    L.vit_decode_f32.restype = ci
    L.vit_decode_f32.argtypes = [vp, vp, vp, vp, ci, ci, ci, vp, sz, vp, vp, vp]
    L.vit_decode_f32_ex.restype = ci
    L.vit_decode_f32_ex.argtypes = [vp, vp, vp, vp, ci, ci, ci, vp, sz, vp, vp, ctypes.POINTER(DecodeOpts), vp]
    L.vit_fb_workspace_bytes.restype = ci
    L.vit_fb_workspace_bytes.argtypes = [ci, ci, ci, ctypes.POINTER(sz)]
    L.vit_forward_backward_f32.restype = ci
    L.vit_forward_backward_f32.argtypes = [vp, vp, vp, vp, ci, ci, ci, vp, sz, vp, vp, vp]
    L.vit_forward_backward_f32_ex.restype = ci
    L.vit_forward_backward_f32_ex.argtypes = [vp, vp, vp, vp, ci, ci, ci, vp, sz, vp, vp, ctypes.POINTER(FbOpts), vp]
    L.vit_analyze_structure_f32.restype = ci
    L.vit_analyze_structure_f32.argtypes = [vp, ci, ctypes.POINTER(Structure)]
    L.vit_clips_in_flight.restype = ci
    L.vit_clips_in_flight.argtypes = [ci, ci, ctypes.POINTER(Structure), ctypes.POINTER(ci)]
    L.vit_emissions_f32.restype = ci
    L.vit_emissions_f32.argtypes = [vp, vp, ci, ci, ci, ci, ci, ctypes.c_float, ci, vp, vp]
    L.vit_voiced_bins.restype = ci
    L.vit_voiced_bins.argtypes = [vp, ctypes.c_longlong, ci, vp, vp, vp]
    L.vit_melody_stats_f32.restype = ci
    L.vit_melody_stats_f32.argtypes = [vp, ci, ci, vp, vp, vp, vp, ci, ci, ci, ctypes.c_float, ctypes.c_float, vp, vp, vp]
    L.vit_upload_frames_f32.restype = ci
    L.vit_upload_frames_f32.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp]
    _lib = L
    return L


def check(code):
    if code != VIT_OK:
        L = load()
        msg = L.vit_strerror(code).decode()
        if code == -5:
            msg += ' [' + L.vit_last_cuda_error().decode() + ']'
        raise VitError(code, msg)


def workspace_bytes(B, T_max, S, algo=ALGO_AUTO):
    out = ctypes.c_size_t(0)
    check(load().vit_workspace_bytes(int(B), int(T_max), int(S), int(algo), ctypes.byref(out)))
    return int(out.value)


def clips_in_flight(S, algo=ALGO_AUTO, structure=None):
    """vit_clips_in_flight: clips one launch keeps in flight with every SM busy (needs a CUDA device)."""
    out = ctypes.c_int(0)
    st = ctypes.byref(structure) if structure is not None else None
    check(load().vit_clips_in_flight(int(S), int(algo), st, ctypes.byref(out)))
    return int(out.value)


def analyze_structure(logA_T):
    """vit_analyze_structure_f32 on a host float32 [S, S] array; returns a Structure."""
    import numpy as np
    A = np.ascontiguousarray(logA_T, np.float32)
    st = Structure()
    check(load().vit_analyze_structure_f32(A.ctypes.data_as(ctypes.c_void_p), int(A.shape[0]), ctypes.byref(st)))
    return st


def launch_count():
    return int(load().vit_launch_count())
