"""vit_analyze_structure_f32 (host-only entry point of the C ABI: runs without a GPU) on the reference's state sets and
on hand-made matrices."""
import numpy as np

from viterbi_spl_b200 import hmm_params


def analyze(cuda_lib, A):
    from viterbi_spl_b200 import _lib
    st = _lib.analyze_structure(A)
    return st.kind, st.halfwidth, st.dense_index, st.background


def test_reference_state_sets(cuda_lib):
    log_tiny = np.float32(np.log(np.finfo(np.float32).tiny))
    for name, S, d in (('dcnet', 321, 12), ('tonet', 361, 14)):
        A, pi = hmm_params.synthetic_hmm(name)
        logA_T, _ = hmm_params.log_params(A, pi)
        kind, hw, di, bg = analyze(cuda_lib, logA_T)
        assert (kind, hw, di) == (1, d, S - 1) and bg == log_tiny
    # jdc: band +-40 of 722 states -- the tensor-memory variant of the banded kernel takes it
    A, pi = hmm_params.synthetic_hmm('jdc')
    kind, hw, di, bg = analyze(cuda_lib, hmm_params.log_params(A, pi)[0])
    assert kind == 1 and hw == 40 and di == 721
    A, pi = hmm_params.synthetic_hmm('imm_hmm')             # band +-56: wide kernel with a shared-memory band tail
    kind, hw, di, bg = analyze(cuda_lib, hmm_params.log_params(A, pi)[0])
    assert kind == 1 and hw == 56 and di == 721
    # imm: fully dense
    A, pi = hmm_params.synthetic_hmm('imm')
    kind, hw, di, bg = analyze(cuda_lib, hmm_params.log_params(A, pi, add_tiny=False)[0])
    assert kind == 0


def test_hand_made_matrices(cuda_lib):
    S = 40
    A = np.full((S, S), -50.0, np.float32)
    for k in range(-3, 4):
        idx = np.arange(max(0, -k), min(S, S - k))
        A[idx, idx + k] = -1.0 - abs(k)
    assert analyze(cuda_lib, A) == (1, 3, -1, -50.0)
    B = A.copy()
    B[7, :] = -2.0
    B[:, 7] = -3.0
    assert analyze(cuda_lib, B) == (1, 3, 7, -50.0)
    C = A.copy()
    C[0, S - 1] = -1.0                                     # one far entry, not a dense state: the band must cover it
    assert analyze(cuda_lib, C)[:2] == (1, S - 1)          # (39 <= 56: the wide kernel's band spans the whole 40 x 40 matrix)
    S3 = 100
    C3 = np.full((S3, S3), -50.0, np.float32)
    for k in (-1, 0, 1):
        idx = np.arange(max(0, -k), min(S3, S3 - k))
        C3[idx, idx + k] = -1.0
    C3[0, S3 - 1] = -1.0                                   # (on a bare diagonal state 0 would be taken as the dense state)
    assert analyze(cuda_lib, C3)[:2] == (0, S3 - 1)        # a band of 99 is no structure at all
    D = np.full((S, S), -np.inf, np.float32)
    D[np.arange(S), np.arange(S)] = 0
    assert analyze(cuda_lib, D) == (1, 0, -1, -np.inf)
    E = A.copy()
    E[3, 3] = np.nan
    assert analyze(cuda_lib, E)[0] == 0
    # band of 57: too wide for either kernel; band of 20 on an odd state count: the wide kernel wants even S
    for S2, d2, want in ((200, 57, 0), (200, 41, 1), (200, 20, 1), (201, 20, 0), (201, 14, 1)):
        F = np.full((S2, S2), -9.0, np.float32)
        for k in range(-d2, d2 + 1):
            idx = np.arange(max(0, -k), min(S2, S2 - k))
            F[idx, idx + k] = -1.0
        assert analyze(cuda_lib, F)[:2] == (want, d2), (S2, d2)
