"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/viterbi_oracle.c (the C restatement of
imm/tf_viterbi.py:75-109).  Used as the checker for cases too large for the NumPy restatement and as the
multi-threaded CPU baseline of bench.py; never imported by the product package."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'libvit_oracle.so')
_lib = None


def build(force=False):
    """Compile viterbi_oracle.c with the committed Makefile (gcc, no fast-math)."""
    src = os.path.join(_HERE, 'viterbi_oracle.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(['make', '-C', _HERE, '-B'], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        f32p = ctypes.POINTER(ctypes.c_float)
        i32p = ctypes.POINTER(ctypes.c_int32)
        i64p = ctypes.POINTER(ctypes.c_int64)
        L.vit_oracle_decode_f32.restype = ctypes.c_int
        L.vit_oracle_decode_f32.argtypes = [f32p, f32p, f32p, ctypes.c_int, ctypes.c_int, i64p, f32p, f32p, i32p]
        L.vit_oracle_decode_batch_f32.restype = ctypes.c_int
        L.vit_oracle_decode_batch_f32.argtypes = [f32p, f32p, f32p, i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                  i64p, f32p, ctypes.c_int]
        L.vit_oracle_max_threads.restype = ctypes.c_int
        L.vit_oracle_decode_long_f32.restype = ctypes.c_int
        L.vit_oracle_decode_long_f32.argtypes = [f32p, f32p, f32p, ctypes.c_int, ctypes.c_int, i64p, f32p, ctypes.c_int]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def viterbi_log_c(logA_T, log_pi, log_emis_ts, return_tables=False):
    """One clip, emissions [T, S].  Returns (states int64[T], score float32[, T1, T2])."""
    A = np.require(logA_T, np.float32, ['C'])
    pi = np.require(log_pi, np.float32, ['C'])
    E = np.require(log_emis_ts, np.float32, ['C'])
    T, S = E.shape
    assert A.shape == (S, S) and pi.shape == (S,)
    states = np.empty([T], np.int64)
    score = np.zeros([1], np.float32)
    T1 = np.empty([T, S], np.float32) if return_tables else None
    T2 = np.empty([T, S], np.int32) if return_tables else None
    rc = lib().vit_oracle_decode_f32(_p(A, ctypes.c_float), _p(pi, ctypes.c_float), _p(E, ctypes.c_float), T, S,
                                     _p(states, ctypes.c_int64), _p(score, ctypes.c_float),
                                     _p(T1, ctypes.c_float) if return_tables else None,
                                     _p(T2, ctypes.c_int32) if return_tables else None)
    assert rc == 0, rc
    if return_tables:
        return states, score[0], T1, T2
    return states, score[0]


def decode_batch_c(logA_T, log_pi, log_emis, lengths=None, nthreads=0):
    """Batch [B, T_max, S] -> (paths int64 [B, T_max] (-1 past length), scores float32 [B])."""
    A = np.require(logA_T, np.float32, ['C'])
    pi = np.require(log_pi, np.float32, ['C'])
    E = np.require(log_emis, np.float32, ['C'])
    B, T_max, S = E.shape
    paths = np.empty([B, T_max], np.int64)
    scores = np.empty([B], np.float32)
    L = None if lengths is None else np.require(lengths, np.int32, ['C'])
    rc = lib().vit_oracle_decode_batch_f32(_p(A, ctypes.c_float), _p(pi, ctypes.c_float), _p(E, ctypes.c_float),
                                           None if L is None else _p(L, ctypes.c_int32), B, T_max, S,
                                           _p(paths, ctypes.c_int64), _p(scores, ctypes.c_float), int(nthreads))
    assert rc == 0, rc
    return paths, scores


def viterbi_log_long_c(logA_T, log_pi, log_emis_ts, nthreads=0):
    """One LONG clip, emissions [T, S], the targets of every step split over `nthreads` threads (0 = all cores);
    bit-identical to viterbi_log_c.  Returns (states int64[T], score float32)."""
    A = np.require(logA_T, np.float32, ['C'])
    pi = np.require(log_pi, np.float32, ['C'])
    E = np.require(log_emis_ts, np.float32, ['C'])
    T, S = E.shape
    assert A.shape == (S, S) and pi.shape == (S,)
    states = np.empty([T], np.int64)
    score = np.zeros([1], np.float32)
    rc = lib().vit_oracle_decode_long_f32(_p(A, ctypes.c_float), _p(pi, ctypes.c_float), _p(E, ctypes.c_float), T, S,
                                          _p(states, ctypes.c_int64), _p(score, ctypes.c_float), int(nthreads))
    assert rc == 0, rc
    return states, score[0]


def max_threads():
    return int(lib().vit_oracle_max_threads())
