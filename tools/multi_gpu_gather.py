"""Multi-GPU plumbing on real GPUs (run under torchrun, one rank per GPU): contiguous clip shards, each rank decodes its
own shard with the CUDA decoder, no data-path collective, final NCCL gather of paths/scores to rank 0, compared with the
CPU oracle there.   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_gpu_gather.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from viterbi_spl_b200 import ViterbiDecoder, hmm_params, sharding, synth

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
A, pi = hmm_params.synthetic_hmm('tonet')
logA_T, log_pi = hmm_params.log_params(A, pi)
B, T, S = 61, 200, 361
E = synth.batch('dense_softmax', B, T, S, seed0=77)                 # every rank builds the same batch, decodes its shard
L = (np.arange(B) * 7 % (T + 1)).astype(np.int32)
dec = ViterbiDecoder(logA_T, log_pi, device=dev)
dE, dL = torch.as_tensor(E).to(dev), torch.as_tensor(L).to(dev)
paths, scores = sharding.decode_sharded(lambda e, l: dec.decode_device(e.contiguous(), l), dE, dL, rank, world, gather=True)
ok = None
if rank == 0:
    from oracle import c_oracle
    want_p, want_s = c_oracle.decode_batch_c(logA_T, log_pi, E, L)
    ok = bool(np.array_equal(paths.cpu().numpy(), want_p) and np.array_equal(scores.cpu().numpy(), want_s))
    sys.stderr.flush()
    print(json.dumps({'world_size': world, 'clips': B, 'gathered_equals_oracle': ok,
                      'shards': [sharding.shard_bounds(B, r, world) for r in range(world)]}))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if (ok is None or ok) else 1)
