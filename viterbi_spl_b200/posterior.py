"""Scaled sum-product forward-backward (posterior state marginals) on the GPU -- the host side above
``vit_forward_backward_f32`` (include/vit_b200.h).

The reference has no such pass (its ``SoftMaxViterbi`` classes are max-product decoders,
dcnet/softmax_viterbi.py:2488-2674); this module runs the textbook scaled recursion on the quantities those classes
hold: the row-stochastic transition matrix and initial distribution loaded from ``viterbi_transition_matrix.dat`` /
``viterbi_init_probs.dat`` (dcnet/softmax_viterbi.py:2581-2618) and the emission likelihoods ``[T, S]`` that
``SoftMaxViterbi.observation_probs_fn`` produces (:2530-2579, probability domain, not logged).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .decoder import checked_lengths


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class ForwardBackward:
    """gamma_t[j] = P(state_t = j | all frames) and log L for batches of clips; float32 on the GPU."""

    def __init__(self, transition_matrix, init_probs, device=None, impl='auto'):
        """impl: 'auto' (the banded kernel when the matrix is band + one dense state with exact zeros elsewhere -- every
        matrix the reference's builders produce -- else the tcgen05 kernel, else the dense FFMA kernel), or 'banded' /
        'tc' / 'simt' to force one (VitError if the shape or the matrix does not allow it)."""
        if not torch.cuda.is_available():
            raise RuntimeError('viterbi_spl_b200 needs a CUDA device (B200); there is no CPU fallback')
        self.lib = _lib.load()
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        A = torch.as_tensor(np.array(transition_matrix, copy=True)) if not torch.is_tensor(transition_matrix) else transition_matrix
        pi = torch.as_tensor(np.array(init_probs, copy=True)) if not torch.is_tensor(init_probs) else init_probs
        assert A.dtype == torch.float32 and pi.dtype == torch.float32, 'parameters must be float32'
        assert A.ndim == 2 and A.shape[0] == A.shape[1] and pi.shape == (A.shape[0],)
        assert bool(torch.all(A >= 0)) and bool(torch.all(pi >= 0))
        assert bool(torch.allclose(A.sum(1), torch.ones(A.shape[0]), atol=1e-4)), 'rows of the transition matrix must sum to 1'
        self.S = int(A.shape[0])
        self.A = A.to(self.device).contiguous()
        self.pi = pi.to(self.device).contiguous()
        self._ws = None
        # band + dense-state structure of the (probability-domain) matrix, found once on the host
        self.structure = _lib.analyze_structure(A.detach().cpu().numpy())
        self.impl = _lib.FB_IMPLS[impl]
        self._opts = _lib.FbOpts()
        self._opts.impl = self.impl
        self._opts.structure = ctypes.pointer(self.structure)

    @property
    def structured(self):
        """True when 'auto' takes the structured kernels for this matrix (unless VIT_FB_IMPL overrides it): band + one
        dense state with exact zeros elsewhere, S <= 384 with half-width <= 14 or S <= 768 with half-width <= 56."""
        st = self.structure
        if not (st.kind == 1 and st.background == 0.0 and self.S >= 2):
            return False
        return (self.S <= 384 and st.halfwidth <= 14) or (self.S <= 768 and st.halfwidth <= 56)

    def run_device(self, lik, lengths=None, gamma=None, loglik=None):
        """lik: CUDA float32 [B, T, S] likelihoods (>= 0); lengths: CUDA int32 [B] or None.
        Returns (gamma float32 [B, T, S], loglik float32 [B]); asynchronous on torch's current stream."""
        assert lik.is_cuda and lik.dtype == torch.float32 and lik.is_contiguous() and lik.ndim == 3
        B, T, S = lik.shape
        assert S == self.S
        if lengths is not None:
            assert lengths.is_cuda and lengths.dtype == torch.int32 and lengths.shape == (B,)
        with torch.cuda.device(self.device):
            n = ctypes.c_size_t(0)
            _lib.check(self.lib.vit_fb_workspace_bytes(B, T, S, ctypes.byref(n)))
            if self._ws is None or self._ws.numel() < n.value:
                self._ws = None
                self._ws = torch.empty(max(n.value, 256), dtype=torch.uint8, device=self.device)
            if gamma is None:
                gamma = torch.empty((B, T, S), dtype=torch.float32, device=self.device)
            if loglik is None:
                loglik = torch.empty((B,), dtype=torch.float32, device=self.device)
            st = torch.cuda.current_stream()
            _lib.check(self.lib.vit_forward_backward_f32_ex(_ptr(self.A), _ptr(self.pi), _ptr(lik), _ptr(lengths), B, T, S,
                                                            _ptr(self._ws), self._ws.numel(), _ptr(gamma), _ptr(loglik),
                                                            ctypes.byref(self._opts), ctypes.c_void_p(st.cuda_stream)))
        return gamma, loglik

    def run_host(self, lik, lengths=None):
        """NumPy in, NumPy out."""
        E = torch.as_tensor(np.ascontiguousarray(lik, np.float32)).to(self.device)
        dL = None if lengths is None else torch.as_tensor(checked_lengths(lengths, E.shape[0], E.shape[1])).to(self.device)
        g, ll = self.run_device(E, dL)
        return g.cpu().numpy(), ll.cpu().numpy()
