"""Batched B200 Viterbi decoder: the host side above the C ABI (include/vit_b200.h).

``decode_batch`` is the new batched entry point; the reference's own entry points
(``viterbi_spl_b200.reference_api``) are thin B=1 wrappers over it.  PyTorch is used only for device memory,
streams and pinned staging buffers; all compute happens in libvit_b200.so.

Tensor layouts (north star): posteriors/emissions ``[B, T, S]`` (``[T, S]`` per clip -- the reference's ``probs`` after
the transpose at imm/tf_viterbi.py:89), initial log-probabilities ``[S]``, transposed log transition matrix
``[S, S]`` dst-major (the reference's ``B = log_transition_matrix_T``, imm/tf_viterbi.py:77) -> int64 state paths
``[B, T]`` (frames past a clip's length are -1) and float32 path scores ``[B]``.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def checked_lengths(lengths, B, T):
    """Host-side clip lengths -> int32 [B], each in [0, T].  The kernels trust the lengths they are given (a device
    pointer cannot be validated by the C ABI): a larger value would read emissions and write the delta history past the
    clip's rows, so every path that still has the lengths on the host checks them here."""
    L = np.asarray(lengths)
    if L.shape != (B,):
        raise ValueError(f'lengths must have shape ({B},), got {L.shape}')
    if L.size and (L.min() < 0 or L.max() > T):
        raise ValueError(f'clip lengths must be in [0, {T}]; got min {L.min()}, max {L.max()}')
    return np.ascontiguousarray(L, np.int32)


class ViterbiDecoder:
    """Holds the HMM parameters on one GPU plus a reusable workspace; decodes batches of clips.

    Parameters are LOG-domain float32, exactly the arrays the reference's decoder objects keep
    (``log_transition_matrix_T`` / ``log_ini_probs``, tonet/softmax_priors.py:1708-1709, 1788-1823).
    """

    def __init__(self, log_transition_matrix_T, log_prob_init, device=None, algo='auto'):
        if not torch.cuda.is_available():
            raise RuntimeError('viterbi_spl_b200 needs a CUDA device (B200); there is no CPU fallback')
        self.lib = _lib.load()
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        # (copies: the reference marks its parameter arrays read-only, which torch cannot wrap)
        A = log_transition_matrix_T if torch.is_tensor(log_transition_matrix_T) else \
            torch.from_numpy(np.array(log_transition_matrix_T, copy=True))
        pi = log_prob_init if torch.is_tensor(log_prob_init) else torch.from_numpy(np.array(log_prob_init, copy=True))
        assert A.dtype == torch.float32 and pi.dtype == torch.float32, 'log-domain parameters must be float32'
        assert A.ndim == 2 and A.shape[0] == A.shape[1], 'log_transition_matrix_T must be [S, S]'
        self.S = int(A.shape[0])
        assert pi.shape == (self.S,), 'log_prob_init must be [S]'
        assert not bool(torch.isnan(A).any()) and not bool(torch.isnan(pi).any()), 'parameters contain NaN'
        # band + background structure of the matrix (every matrix the reference's builders produce has it): unlocks the
        # bit-exact VIT_ALGO_BANDED fast path under algo='auto'
        self.structure = _lib.analyze_structure(A.detach().cpu().numpy())
        self.logA_T = A.to(self.device).contiguous()
        self.log_pi = pi.to(self.device).contiguous()
        self.algo = _lib.ALGO_NAMES[algo] if isinstance(algo, str) else int(algo)
        self._ws = None
        self._pinned = {}
        self._dev_emis = None
        self._copy_stream = None

    # ---- device path ---------------------------------------------------------------------------------------
    def _workspace(self, nbytes, slot=0):
        if self._ws is None:
            self._ws = {}
        ws = self._ws.get(slot)
        if ws is None or ws.numel() < nbytes:
            self._ws[slot] = None
            ws = self._ws[slot] = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
        return ws

    def decode_device(self, log_emis, lengths=None, paths=None, scores=None, want_tables=False, stream=None,
                      forward_events=None, frame_range=None, backtrace=True, backtrace_stream=None, workspace_slot=0):
        """log_emis: CUDA float32 [B, T, S] contiguous; lengths: CUDA int32 [B] or None.
        Returns (paths int64 [B, T], scores float32 [B]) CUDA tensors (+ (T1 float32, T2 uint16) if want_tables).
        Asynchronous on `stream` (default: torch's current stream).  forward_events: optional pair of
        torch.cuda.Event(enable_timing=True) recorded by the library around the forward kernel.
        frame_range=(t0, t1) runs the recursion over those frames only, resuming (t0 > 0) from the delta history the
        previous call left in this decoder's workspace; backtrace=False skips the backtrace (all but the last range).
        Frame ranges need the tmem or the banded algorithm.
        backtrace_stream: run the backtrace on that stream (ordered after the forward kernel) so that the next decode
        on `stream` overlaps it; the caller then owns the synchronisation (see PipelinedDecoder) and must give
        consecutive calls different `workspace_slot`s."""
        assert log_emis.is_cuda and log_emis.dtype == torch.float32 and log_emis.is_contiguous()
        B, T, S = log_emis.shape
        assert S == self.S, f'emissions have {S} states, model has {self.S}'
        if lengths is not None:
            assert lengths.is_cuda and lengths.dtype == torch.int32 and lengths.shape == (B,)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream() if stream is None else stream
            with torch.cuda.stream(st):
                if paths is None:
                    paths = torch.empty((B, T), dtype=torch.int64, device=self.device)
                if scores is None:
                    scores = torch.empty((B,), dtype=torch.float32, device=self.device)
                algo = _lib.ALGO_BACKPOINTER if want_tables else self.algo
                ws = self._workspace(_lib.workspace_bytes(B, T, S, algo), workspace_slot)   # (banded needs no more than auto's choice)
                opts = _lib.DecodeOpts(algo=algo)
                T1 = T2 = None
                if want_tables:
                    T1 = torch.zeros((B, T, S), dtype=torch.float32, device=self.device)
                    T2 = torch.zeros((B, T, S), dtype=torch.uint16, device=self.device)
                    opts.d_delta = T1.data_ptr()
                    opts.d_backpointers = T2.data_ptr()
                if forward_events is not None:
                    for ev in forward_events:   # torch creates the cudaEvent lazily on first record
                        if ev.cuda_event == 0:
                            ev.record(st)
                    opts.ev_forward_begin = forward_events[0].cuda_event
                    opts.ev_forward_end = forward_events[1].cuda_event
                if frame_range is not None:
                    opts.frame_begin, opts.frame_end = int(frame_range[0]), int(frame_range[1])
                    assert 0 <= opts.frame_begin < opts.frame_end <= T, 'frame_range must be a non-empty range in [0, T]'
                opts.skip_backtrace = 0 if backtrace else 1
                if backtrace_stream is not None:
                    opts.backtrace_stream = backtrace_stream.cuda_stream
                if algo in (_lib.ALGO_AUTO, _lib.ALGO_BANDED):
                    opts.structure = ctypes.pointer(self.structure)
                rc = self.lib.vit_decode_f32_ex(_ptr(self.logA_T), _ptr(self.log_pi), _ptr(log_emis), _ptr(lengths),
                                                B, T, S, _ptr(ws), ws.numel(), _ptr(paths), _ptr(scores),
                                                ctypes.byref(opts), ctypes.c_void_p(st.cuda_stream))
                _lib.check(rc)
        if want_tables:
            return paths, scores, T1, T2
        return paths, scores

    # ---- host path (what the reference-facing wrappers use) --------------------------------------------------
    def _pinned_buf(self, key, shape, dtype):
        n = int(np.prod(shape))
        buf = self._pinned.get(key)
        if buf is None or buf.numel() < n or buf.dtype != dtype:
            buf = torch.empty(max(n, 1), dtype=dtype).pin_memory()
            self._pinned[key] = buf
        return buf[:n].view(*shape)

    def supports_frame_slabs(self):
        """True if this decoder's algorithm can run the recursion as consecutive frame ranges (tmem, banded)."""
        algo = self.algo
        if algo == _lib.ALGO_AUTO:
            algo = self.lib.vit_select_algo(1, 1, self.S)      # (auto takes the banded kernel where the matrix allows)
        if algo == _lib.ALGO_BANDED:
            return self.structure.kind == 1
        return algo in (_lib.ALGO_TMEM, _lib.ALGO_STREAM)

    def decode_host(self, log_emis, lengths=None, slab_frames=None, out=None):
        """log_emis: host float32 array [B, T, S] (NumPy or CPU tensor).  Copies host->device, decodes, copies the
        paths and scores back; returns NumPy (paths int64 [B, T], scores float32 [B]).

        Large batches are uploaded in time slabs of `slab_frames` frames (default: about T/16 once the batch exceeds
        64 MB) on a copy stream while the recursion over the previous slab runs on the compute stream, so the
        end-to-end time is max(PCIe copy, decode) instead of their sum.  (Measured on B200: 16 slabs of a 4.4 GB batch
        upload at the full 55 GB/s; 32 slabs -- 136 KB rows -- drop to 35 GB/s.)
        out=(paths, scores): optional page-locked CPU tensors (int64 [B, T], float32 [B]) that receive the results
        directly; they are returned as NumPy views, which saves a 25 MB host copy per 1024 x 3000 batch."""
        E = torch.as_tensor(log_emis)
        assert E.dtype == torch.float32 and E.ndim == 3
        B, T, S = E.shape
        with torch.cuda.device(self.device):
            if E.is_pinned() and E.is_contiguous():
                src = E
            else:
                src = self._pinned_buf('emis', (B, T, S), torch.float32)
                src.copy_(E)
            dL = None
            if lengths is not None:
                dL = torch.as_tensor(checked_lengths(lengths, B, T)).to(self.device)
            if slab_frames is None:
                slab_frames = T if B * T * S * 4 < (64 << 20) else max(64, -(-T // 16))
            if slab_frames >= T or B == 0 or not self.supports_frame_slabs():
                dE = src.to(self.device, non_blocking=True)
                paths, scores = self.decode_device(dE, dL)
            else:
                paths, scores = self._decode_host_slabs(src, dL, int(slab_frames))
            if out is not None:
                hp, hs = out
                assert hp.is_pinned() and hs.is_pinned() and hp.shape == (B, T) and hs.shape == (B,)
                assert hp.dtype == torch.int64 and hs.dtype == torch.float32
            else:
                hp = self._pinned_buf('paths', (B, T), torch.int64)
                hs = self._pinned_buf('scores', (B,), torch.float32)
            hp.copy_(paths, non_blocking=True)
            hs.copy_(scores, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        if out is not None:
            return hp.numpy(), hs.numpy()
        return hp.numpy().copy(), hs.numpy().copy()

    def decode_host_st(self, log_emis_st):
        """ONE recording given as the reference gives it to its log-domain entry point: host float32 ``[S, T]``
        (imm/tf_viterbi.py:75-89).  Uploaded as it is and transposed to ``[T, S]`` on the device -- the reference's
        host-side transposing copy is the most expensive thing left in a single-recording call.  Returns NumPy
        (path int64 [T], score float32)."""
        E = torch.as_tensor(log_emis_st)
        assert E.dtype == torch.float32 and E.ndim == 2 and E.shape[0] == self.S
        S, T = E.shape
        with torch.cuda.device(self.device):
            src = self._pinned_buf('emis_st', (S, T), torch.float32)
            src.copy_(E)
            dE = src.to(self.device, non_blocking=True).t().contiguous().view(1, T, S)
            paths, scores = self.decode_device(dE)
            hp = self._pinned_buf('paths', (1, T), torch.int64)
            hs = self._pinned_buf('scores', (1,), torch.float32)
            hp.copy_(paths, non_blocking=True)
            hs.copy_(scores, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return hp.numpy()[0].copy(), hs.numpy()[0].copy()

    def _decode_host_slabs(self, src, dL, slab):
        """Upload frames [t0, t1) of every clip (one strided 2-D copy per slab) on the copy stream; the compute
        stream runs the recursion over slab k as soon as its copy has landed and while slab k+1 is in flight."""
        B, T, S = src.shape
        n = B * T * S
        if self._dev_emis is None or self._dev_emis.numel() < n:
            self._dev_emis = None
            self._dev_emis = torch.empty(n, dtype=torch.float32, device=self.device)
        dE = self._dev_emis[:n].view(B, T, S)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        main, cp = torch.cuda.current_stream(), self._copy_stream
        cp.wait_stream(main)                     # the previous call's kernels may still read the device buffer
        paths = torch.empty((B, T), dtype=torch.int64, device=self.device)
        scores = torch.empty((B,), dtype=torch.float32, device=self.device)
        for t0 in range(0, T, slab):
            t1 = min(T, t0 + slab)
            _lib.check(self.lib.vit_upload_frames_f32(_ptr(dE), ctypes.c_void_p(src.data_ptr()), B, T, S, t0, t1,
                                                      ctypes.c_void_p(cp.cuda_stream)))
            ev = torch.cuda.Event()
            ev.record(cp)
            main.wait_event(ev)
            self.decode_device(dE, dL, paths, scores, frame_range=(t0, t1), backtrace=(t1 == T))
        return paths, scores

    def decode(self, log_emis, lengths=None):
        """Dispatch on where the emissions live: CUDA tensor -> CUDA tensors, host array -> NumPy arrays."""
        if torch.is_tensor(log_emis) and log_emis.is_cuda:
            return self.decode_device(log_emis, lengths)
        return self.decode_host(log_emis, lengths)


class PipelinedDecoder:
    """Back-to-back batches with the (latency-bound) backtrace of batch k overlapping the forward recursion of batch
    k+1: the forward kernels run on torch's current stream, the backtraces on a second stream, and two workspaces
    alternate so that a history is never overwritten while it is still being walked.

        pd = PipelinedDecoder(decoder)
        for emis in batches:                       # CUDA tensors [B, T, S]
            paths, scores = pd.submit(emis)        # asynchronous; results valid after pd.finish() (or pd.wait(k))
        pd.finish()
    """

    def __init__(self, decoder):
        self.dec = decoder
        self.bt_stream = torch.cuda.Stream(device=decoder.device)
        self._done = [None, None]                  # event: backtrace that last used workspace slot i has finished
        self._k = 0

    def submit(self, log_emis, lengths=None, paths=None, scores=None, forward_events=None):
        slot = self._k & 1
        main = torch.cuda.current_stream()
        if self._done[slot] is not None:
            main.wait_event(self._done[slot])      # the forward about to run overwrites that slot's history
        out = self.dec.decode_device(log_emis, lengths, paths, scores, forward_events=forward_events,
                                     backtrace_stream=self.bt_stream, workspace_slot=slot)
        # the backtrace reads the lengths and the workspace and writes paths / scores on bt_stream: tell the caching
        # allocator, or a block the caller drops before finish() could be handed out while the walk still uses it
        for t in (out[0], out[1], lengths, self.dec._ws.get(slot)):
            if t is not None:
                t.record_stream(self.bt_stream)
        ev = torch.cuda.Event()
        ev.record(self.bt_stream)
        self._done[slot] = ev
        self._k += 1
        return out

    def finish(self):
        """Make torch's current stream wait for every outstanding backtrace (results are then ordered as usual)."""
        main = torch.cuda.current_stream()
        for ev in self._done:
            if ev is not None:
                main.wait_event(ev)


def decode_batch(log_emis, logA_T, log_pi, lengths=None, algo='auto'):
    """One-shot batched decode: ``log_emis [B, T, S]``, ``logA_T [S, S]`` (dst-major), ``log_pi [S]``,
    ``lengths [B]`` or None -> ``(paths int64 [B, T], scores float32 [B])``.  Inputs on the GPU give GPU outputs."""
    device = log_emis.device if (torch.is_tensor(log_emis) and log_emis.is_cuda) else None
    return ViterbiDecoder(logA_T, log_pi, device=device, algo=algo).decode(log_emis, lengths)
