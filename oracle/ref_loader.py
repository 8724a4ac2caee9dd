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


class _NPTensor(np.ndarray):
    """ndarray that answers the two tf.Tensor methods the reference's peak finder calls on its intermediate values."""

    def numpy(self):
        return np.asarray(self)

    def set_shape(self, shape):
        assert len(shape) == self.ndim and all(a is None or a == b for a, b in zip(shape, self.shape))


class _FakeSignal:
    @staticmethod
    def frame(x, frame_length, frame_step, axis=-1):
        assert frame_step == 1 and axis == -1
        return np.lib.stride_tricks.sliding_window_view(np.asarray(x), frame_length, axis=-1).view(_NPTensor)


class FakeTF:
    """NumPy stand-ins for the handful of TensorFlow calls inside ``Viterbi.find_peaks_all_at_once_tf_fn``
    (dcnet/softmax_viterbi.py:2294-2314: convert_to_tensor, pad(reflect), signal.frame, argmax) so that the reference's
    own ``observation_probs_fn`` / ``__call__`` of the dcnet and msnet ``class Viterbi`` can be EXECUTED here without
    TensorFlow.  ``tf.argmax`` documents no tie order; NumPy's first maximum is the pinned behaviour."""
    int32 = np.int32
    float32 = np.float32
    Tensor = _NPTensor
    Variable = _FakeTFVar
    signal = _FakeSignal

    @staticmethod
    def function(*a, **kw):
        return (lambda f: f) if not (a and callable(a[0])) else a[0]

    @staticmethod
    def TensorSpec(*a, **kw):
        return None

    @staticmethod
    def convert_to_tensor(x, dtype=None):
        return np.asarray(x, dtype).view(_NPTensor)

    @staticmethod
    def pad(x, paddings, mode='constant'):
        return np.pad(np.asarray(x), paddings, mode=mode.lower()).view(_NPTensor)

    @staticmethod
    def argmax(x, axis=-1, output_type=np.int64):
        return np.argmax(np.asarray(x), axis=axis).astype(output_type).view(_NPTensor)


class _NPMath:
    @staticmethod
    def count_nonzero(x, dtype=np.int64):
        return np.asarray(np.count_nonzero(np.asarray(x)), dtype).view(_NPTensor)


def _np_tensor(x, dtype=None):
    return np.asarray(x, dtype).view(_NPTensor)


class FakeTFStats(FakeTF):
    """FakeTF plus the float32 element-wise / reduction ops of ``MetricsInference.viterbi_update_states_tf_fn`` and
    ``MetricsBase.est_notes_fn`` / ``octave`` / ``count_nonzero_fn`` (dcnet/softmax_viterbi.py:1919-1958, 2923-2979), each
    mapped to the NumPy ufunc of the same name on float32 arrays -- enough to EXECUTE those reference functions.
    (``tf.sigmoid`` is 1 / (1 + exp(-x)) in float32 here; TensorFlow's kernel may differ in the last ulp, which is why the
    product is held to 1e-5 on notes and the counters may differ on frames within 2e-5 of a threshold.)"""
    bool = np.bool_
    int64 = np.int64
    math = _NPMath

    convert_to_tensor = staticmethod(lambda x, dtype=None: _np_tensor(x, dtype))
    constant = staticmethod(lambda x, dtype=None: _np_tensor(x, dtype))
    cast = staticmethod(lambda x, dtype: _np_tensor(np.asarray(x).astype(dtype)))
    logical_not = staticmethod(lambda x: np.logical_not(x).view(_NPTensor))
    logical_and = staticmethod(lambda a, b: np.logical_and(a, b).view(_NPTensor))
    abs = staticmethod(lambda x: np.abs(x).view(_NPTensor))
    floor = staticmethod(lambda x: np.floor(x).view(_NPTensor))
    maximum = staticmethod(lambda a, b: np.maximum(a, np.asarray(b, np.asarray(a).dtype)).view(_NPTensor))
    zeros_like = staticmethod(lambda x: np.zeros_like(np.asarray(x)).view(_NPTensor))
    where = staticmethod(lambda c, a, b: np.where(c, a, b).view(_NPTensor))
    reduce_sum = staticmethod(lambda x, axis=None: np.sum(np.asarray(x), axis=axis, dtype=np.asarray(x).dtype).view(_NPTensor))
    range = staticmethod(lambda n, dtype=np.int32: np.arange(n, dtype=dtype).view(_NPTensor))
    size = staticmethod(lambda x, out_type=np.int32: np.asarray(np.asarray(x).size, out_type).view(_NPTensor))

    @staticmethod
    def sigmoid(x):
        x = np.asarray(x, np.float32)
        return (np.float32(1) / (np.float32(1) + np.exp(-x))).astype(np.float32).view(_NPTensor)


def dcnet_melody_stats():
    """(f)3: the reference's OWN ``MetricsInference.viterbi_update_states_tf_fn`` (dcnet/softmax_viterbi.py:2923-2979) with its
    ``MetricsBase`` helpers (:1919-1958), executed on FakeTFStats.  Returns
    ``fn(ref_notes [T] f32, logits [T, 320] f32, melody_bins [T] int32, est_voicing [T] bool) -> (est_notes_with_voicing_info
    [T] f32, {counter name: int})`` with the counters in the names the reference's variable dictionary uses."""
    relpath = 'dcnet/softmax_viterbi.py'
    tf = FakeTFStats
    note_range = (np.arange(320) / 5. + 23.6).astype(np.float32)                 # TFDataset.note_range, :428-432
    ns = {'tf': tf, 'TFDataset': type('TFDataset', (), {'note_range': note_range})}
    base = type('MetricsBase', (), {name: staticmethod(ref_method(relpath, 'MetricsBase', name, extra=ns))
                                    for name in ('count_nonzero_fn', 'est_notes_fn', 'octave')})
    ns['MetricsBase'] = base
    update = ref_method(relpath, 'MetricsInference', 'viterbi_update_states_tf_fn', extra=ns)

    class _Self:
        def __init__(self):
            self.values = {}
            self.viterbi_var_dict = {'all_updated': {0: True}}

        def update_melody_var_fn(self, rec_idx, l1, l2, value, viterbi=False):
            assert viterbi
            self.values[f'{l1}_{l2}'] = int(value)

    def fn(ref_notes, logits, melody_bins, est_voicing):
        me = _Self()
        out = update(me, 0, ref_notes, logits, melody_bins, est_voicing)
        return np.asarray(out, np.float32), me.values
    return fn


# (file, class) of every ``class Viterbi`` / ``class SoftMaxViterbi`` copy the drop-in namespaces mirror
CLASS_COPIES = {
    ('dcnet', 'Viterbi'): 'dcnet/softmax_viterbi.py', ('dcnet', 'SoftMaxViterbi'): 'dcnet/softmax_viterbi.py',
    ('msnet', 'Viterbi'): 'msnet/viterbi_softmax.py', ('msnet', 'SoftMaxViterbi'): 'msnet/viterbi_softmax.py',
    ('ftanet', 'Viterbi'): 'ftanet/viterbi_performance.py', ('ftanet', 'SoftMaxViterbi'): 'ftanet/viterbi_performance.py',
    ('jdc', 'Viterbi'): 'jdc/viterbi_softmax.py', ('jdc', 'SoftMaxViterbi'): 'jdc/viterbi_softmax.py',
    ('tonet', 'Viterbi'): 'tonet/softmax_priors.py', ('tonet', 'SoftMaxViterbi'): 'tonet/softmax_priors.py',
    ('tonet', 'ViterbiA'): 'tonet/main_shaun.py', ('tonet', 'SoftMaxViterbiAlwaysScaled'): 'tonet/ablation.py',
    ('tonet', 'SoftMaxViterbiWide'): 'tonet/for_paper.py',
    ('imm', 'Viterbi'): 'imm/main_imm.py',
}


def reference_class(namespace, name):
    """The reference's own class object for one entry of CLASS_COPIES (AST-exec'd, TensorFlow stubbed by FakeTF)."""
    relpath = CLASS_COPIES[(namespace, name)]
    cls_name = {'ViterbiA': 'Viterbi', 'SoftMaxViterbiAlwaysScaled': 'SoftMaxViterbi',
                'SoftMaxViterbiWide': 'SoftMaxViterbi'}.get(name, name)
    return ref_toplevel(relpath, cls_name, extra={'load_np_array_from_file_fn': dat_loader(), 'tf': FakeTF})


def construct_in(directory, cls, *args, **kw):
    """Run a reference constructor with `directory` as the working directory (they read their .dat files from cwd)."""
    with _Chdir(directory):
        return cls(*args, **kw)


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
