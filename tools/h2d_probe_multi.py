"""Host -> device upload bandwidth with N ranks copying at once (one rank per GPU, as bench.py's end-to-end leg does):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/h2d_probe_multi.py

Every rank pins a 1024 x 3000 x 361 float32 batch (4.44 GB) after binding itself to its GPU's NUMA node, then -- all
ranks between the same two barriers -- uploads it (a) as ONE contiguous cudaMemcpyAsync, (b) as 16 time slabs, each a
strided 2-D copy of [B] rows (what ViterbiDecoder.decode_host issues: vit_upload_frames_f32), (c) as 16 contiguous slabs
of a time-major staging buffer [slab][B][frames][S].  Rank 0 prints one JSON line with the aggregate GB/s (max over
ranks of the elapsed time).  Explains where the 8-GPU end-to-end curve of SCALE_r01.json flattens."""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
numa = bench.bind_to_gpu_numa_node(local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
from viterbi_spl_b200 import _lib
L = _lib.load()
B, T, S = 1024, 3000, 361
host = torch.empty((B, T, S), dtype=torch.float32).pin_memory()
host.zero_()
d = torch.empty((B, T, S), dtype=torch.float32, device=dev)
st = torch.cuda.current_stream()
nbytes = host.numel() * 4


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = float(t[0]) if best is None else min(best, float(t[0]))
    return best


def contiguous():
    d.copy_(host, non_blocking=True)


def strided_slabs(slab=188):
    for a in range(0, T, slab):
        L.vit_upload_frames_f32(ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(host.data_ptr()), B, T, S, a, min(T, a + slab),
                                ctypes.c_void_p(st.cuda_stream))


def contiguous_slabs(slab=188):
    flat_h, flat_d = host.view(-1), d.view(-1)
    n = B * slab * S
    for a in range(0, flat_h.numel(), n):
        flat_d[a:a + n].copy_(flat_h[a:a + n], non_blocking=True)


out = {'n_gpus': world, 'bytes_per_gpu': nbytes, 'numa_cpus_bound': numa}
for name, fn in (('one_contiguous_copy', contiguous), ('16_strided_2d_slabs', strided_slabs), ('16_contiguous_slabs', contiguous_slabs)):
    dt = timed(fn)
    out[name] = {'ms': dt * 1e3, 'aggregate_GBps': world * nbytes / dt / 1e9, 'per_gpu_GBps': nbytes / dt / 1e9}
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
