"""viterbi_spl_b200 -- B200-native batched HMM (Viterbi) decoder for pitch-bin melody models.

A from-scratch sm_100a implementation of the one data-parallel hot path of drwangxian/viterbi_spl: the float32
log-domain Viterbi recursion + backtrace over a dense [S, S] transition matrix, batched over clips.

    from viterbi_spl_b200 import ViterbiDecoder, decode_batch        # batched API (torch CUDA tensors or NumPy)
    from viterbi_spl_b200 import reference_api                       # the reference's own entry points, drop-in

All compute runs in ``libvit_b200.so`` (hand-written CUDA behind the C ABI in include/vit_b200.h); there is no CPU
fallback -- constructing a decoder without the library or without a GPU raises.
"""
from . import _lib, hmm_params, synth  # noqa: F401
from .decoder import PipelinedDecoder, ViterbiDecoder, decode_batch  # noqa: F401
from .posterior import ForwardBackward  # noqa: F401

__version__ = '0.1.0'
