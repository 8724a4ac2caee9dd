"""World-size-2 test of the multi-GPU plumbing on CPU (gloo): contiguous clip shards, no data-path collective, final
gather to rank 0.  The per-rank decode function is injected; here it is the CPU oracle (the CUDA decoder needs a GPU and
is exercised by the gpu-marked tests), so this covers exactly the host logic bench.py --gpus N relies on."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import np_oracle
from viterbi_spl_b200 import sharding, synth


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, T, S, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        A, pi = synth.dyadic_hmm(S, seed=11)
        E = synth.batch('dyadic', B, T, S, seed0=3)
        L = (np.arange(B) % (T + 1)).astype(np.int32)

        def decode_fn(emis, lengths):
            return np_oracle.decode_batch_np(A, pi, emis, lengths)

        lo, hi, p_shard, s_shard = sharding.decode_sharded(decode_fn, E, L, rank, world)
        assert (lo, hi) == sharding.shard_bounds(B, rank, world) and p_shard.shape == (hi - lo, T)
        paths, scores = sharding.decode_sharded(decode_fn, E, L, rank, world, gather=True)
        # timing plumbing of bench.py: max over ranks
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t[0]) == float(world)
        dist.barrier()
        if rank == 0:
            np.savez(os.path.join(out_dir, 'gathered.npz'), paths=paths, scores=scores)
        else:
            assert paths is None and scores is None
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_decode_matches_single_process(tmp_path):
    B, T, S = 7, 12, 23          # odd clip count: ranks get 4 and 3 clips
    mp.spawn(_worker, args=(2, _free_port(), B, T, S, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / 'gathered.npz')
    A, pi = synth.dyadic_hmm(S, seed=11)
    E = synth.batch('dyadic', B, T, S, seed0=3)
    L = (np.arange(B) % (T + 1)).astype(np.int32)
    want_p, want_s = np_oracle.decode_batch_np(A, pi, E, L)
    assert np.array_equal(got['paths'], want_p)
    assert np.array_equal(got['scores'], want_s)
