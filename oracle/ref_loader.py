"""TEST INFRASTRUCTURE ONLY -- executes the reference's OWN decode functions from /root/reference.

The reference's modules import TensorFlow/librosa at import time (self_defined/__init__.py:1-8), which are not
installed, so single ``def``/``class`` nodes are extracted by AST and exec'd with NumPy only.
Nothing is copied into this repository: the source text is read from the read-only checkout at call time.
/root/reference does not exist on the GPU box; callers must check ``available()``.
"""
import ast
import importlib.util
import os
import textwrap

import numpy as np

REF_ROOT = os.environ.get('VITERBI_SPL_REFERENCE', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'imm', 'tf_viterbi.py'))


def _source(relpath):
    with open(os.path.join(REF_ROOT, relpath)) as fh:
        return fh.read()


def ref_toplevel(relpath, name, extra=None):
    """exec ONE top-level def/class of a reference file; returns the object."""
    src = _source(relpath)
    ns = {'np': np, **(extra or {})}
    for node in ast.parse(src).body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name == name:
            seg = ast.get_source_segment(src, node)
            # drop decorators that need TF / numba (e.g. @tf.function on methods is handled by ref_method)
            exec(seg, ns)
            return ns[name]
    raise KeyError(f'{name} not found in {relpath}')


def ref_method(relpath, cls, name, extra=None):
    """exec ONE method of a reference class as a plain function (decorators stripped)."""
    src = _source(relpath)
    ns = {'np': np, **(extra or {})}
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == name:
                    lines = src.splitlines()[sub.lineno - 1:sub.end_lineno]   # from `def`, decorators excluded
                    exec(textwrap.dedent('\n'.join(lines)), ns)
                    return ns[name]
    raise KeyError(f'{cls}.{name} not found in {relpath}')


def dat_loader():
    """The reference's .dat reader (self_defined/load_np_array_from_file.py:4-27), imported by file path."""
    path = os.path.join(REF_ROOT, 'self_defined', 'load_np_array_from_file.py')
    spec = importlib.util.spec_from_file_location('_ref_load_np_array', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.load_np_array_from_file_fn


def dat_saver():
    path = os.path.join(REF_ROOT, 'self_defined', 'save_np_array_to_file.py')
    spec = importlib.util.spec_from_file_location('_ref_save_np_array', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.save_np_array_to_file_fn


# ---- the functions of SURVEY.md section 8(a) that run without TensorFlow -------------------------------------

def log_domain_decode():
    """a1: imm/tf_viterbi.py:75-109 -- the oracle of record."""
    return ref_toplevel('imm/tf_viterbi.py', 'viterbi_librosa_fn')


def family_a_decode(relpath='dcnet/softmax_viterbi.py', cls='Viterbi'):
    """a6: static Family-A decode (prob-domain in), e.g. dcnet/softmax_viterbi.py:2433-2485."""
    return ref_method(relpath, cls, 'viterbi_librosa_fn')


def dcnet_c_decode():
    """a6: dcnet/tf_viterbi_decoding.py:156-207 viterbi_librosa_c_fn."""
    return ref_toplevel('dcnet/tf_viterbi_decoding.py', 'viterbi_librosa_c_fn')


def dcnet_f64_decode():
    """a7: dcnet/tf_viterbi_decoding.py:209-263 (float64 T1 table; not the parity target)."""
    return ref_toplevel('dcnet/tf_viterbi_decoding.py', 'viterbi_librosa_fn')


def numba_core():
    """a4: dcnet/tf_viterbi_decoding.py:75-116, the numba-jit twin of the AOT `viterbi_numba.core`."""
    import numba
    src = _source('dcnet/tf_viterbi_decoding.py')
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == '_viterbi_core_numba_fn':
            lines = src.splitlines()[node.lineno - 1:node.end_lineno]
            ns = {'np': np, 'numba': numba}
            exec('\n'.join(lines), ns)
            sig = numba.int64[:](numba.float32[:, ::1], numba.float32[:], numba.float32[:, ::1])
            return numba.jit(sig, nopython=True)(ns['_viterbi_core_numba_fn'])
    raise KeyError('_viterbi_core_numba_fn')


class _Chdir:
    def __init__(self, path):
        self.path = path

    def __enter__(self):
        self.old = os.getcwd()
        os.chdir(self.path)

    def __exit__(self, *a):
        os.chdir(self.old)


class _FakeTFVar:
    """Stands in for the tf.Variable the dcnet SoftMaxViterbi ctor asserts on (dcnet/softmax_viterbi.py:2497)."""

    def __init__(self, v):
        self.v = np.float32(v)

    def numpy(self):
        return self.v


def msnet_softmax_viterbi(scaled, voicing_threshold=0.5):
    """a9/a11: msnet/viterbi_softmax.py:1732-1910 class SoftMaxViterbi with the SHIPPED msnet/*.dat parameters."""
    cls = ref_toplevel('msnet/viterbi_softmax.py', 'SoftMaxViterbi', extra={'load_np_array_from_file_fn': dat_loader()})
    with _Chdir(os.path.join(REF_ROOT, 'msnet')):
        try:
            return cls(scaled=scaled)
        except TypeError:
            return cls(voicing_threshold, scaled)


def tonet_viterbi_class():
    """a8: tonet/softmax_priors.py:1691-1878 class Viterbi (Family B; needs S=361 .dat files in cwd)."""
    return ref_toplevel('tonet/softmax_priors.py', 'Viterbi', extra={'load_np_array_from_file_fn': dat_loader()})


def tonet_softmax_viterbi_class():
    """a9: tonet/softmax_priors.py:1881-2061 class SoftMaxViterbi (Family C)."""
    return ref_toplevel('tonet/softmax_priors.py', 'SoftMaxViterbi',
                        extra={'load_np_array_from_file_fn': dat_loader()})


def imm_viterbi_class():
    """a10: imm/tf_imm.py:48-135 class Viterbi (Family D) with imm/transition_matrix.py:4-31."""
    gen = ref_toplevel('imm/transition_matrix.py', 'gen_transition_matrix_fn')

    class _TF:  # `isinstance(HF0, (np.ndarray, tf.Tensor))` at imm/tf_imm.py:72
        Tensor = type('Tensor', (), {})

    return ref_toplevel('imm/tf_imm.py', 'Viterbi', extra={'gen_transition_matrix_fn': gen, 'tf': _TF})


def imm_gen_transition_matrix():
    return ref_toplevel('imm/transition_matrix.py', 'gen_transition_matrix_fn')


def run_ref_script(relpath, files):
    """Execute one of the reference's top-level parameter scripts (e.g. dcnet/viterbi_transition_matrix.py) in a temp
    directory holding the given input arrays as `<name>.dat`, with stub modules for its TensorFlow/plotting imports.
    Returns {record name: array} for every .dat the script wrote."""
    import contextlib
    import io
    import runpy
    import sys
    import tempfile
    import types

    saver, loader = dat_saver(), dat_loader()
    sd = types.ModuleType('self_defined')
    sd.load_np_array_from_file_fn = loader
    sd.save_np_array_to_file_fn = saver

    class _Any(types.ModuleType):
        def __getattr__(self, k):
            return lambda *a, **kw: None

    stubs = {'self_defined': sd}
    for m in ['librosa', 'matplotlib', 'matplotlib.pyplot', 'scipy.stats']:
        stubs[m] = _Any(m)
    stubs['matplotlib'].pyplot = stubs['matplotlib.pyplot']
    old = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    cwd = os.getcwd()
    out = {}
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            for name, arr in files.items():
                saver(name + '.dat', arr, name)
            with contextlib.redirect_stdout(io.StringIO()):
                runpy.run_path(os.path.join(REF_ROOT, relpath), run_name='__main__')
            for f in os.listdir(d):
                if f.endswith('.dat') and f[:-4] not in files:
                    out[f[:-4]] = loader(f)[1]
        finally:
            os.chdir(cwd)
            for k, v in old.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
    return out
