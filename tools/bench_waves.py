"""BASELINE.json configs 3 and 5 at their NAMED sizes, decoded in HBM-sized waves (viterbi_spl_b200.waves).

    python tools/bench_waves.py --config 5 [--algo dense|auto]     # 65,536 clips x 3000 frames x 361 states
    python tools/bench_waves.py --config 3 [--algo dense|auto]     # 4096 clips x 10,000 frames x 722 states
    torchrun --nproc-per-node N ... tools/bench_waves.py --config 5   # the clips sharded over N GPUs (strong scaling)

Emissions are generated on the device, wave by wave, on a side stream (seed = first clip of the wave) while the previous
wave is decoded; `decode_ms` sums the CUDA-event time of the decode calls alone, `job_ms` is the whole job including
whatever generation did not hide.  Parity: the first `--check` clips of the first wave are pulled to the host and
decoded by the oracle (test infrastructure) -- paths and scores must be bit-equal.  Prints one JSON line (rank 0).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

CONFIGS = {3: dict(clips=4096, frames=10000, states=722, model='jdc'),
           5: dict(clips=65536, frames=3000, states=361, model='tonet')}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', type=int, default=5, choices=sorted(CONFIGS))
    ap.add_argument('--algo', default='dense')
    ap.add_argument('--clips', type=int, default=None)
    ap.add_argument('--model', default=None, help='jdc (band +-40) or imm (fully dense) for config 3')
    ap.add_argument('--check', type=int, default=2, help='clips of the first wave checked against the oracle')
    ap.add_argument('--max-wave-clips', type=int, default=None)
    a = ap.parse_args()
    cfg = dict(CONFIGS[a.config])
    if a.clips:
        cfg['clips'] = a.clips
    if a.model:
        cfg['model'] = a.model

    import torch.distributed as dist
    from viterbi_spl_b200 import ViterbiDecoder, _lib, hmm_params, synth, sharding
    from viterbi_spl_b200.waves import WaveDecoder

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    T, S = cfg['frames'], cfg['states']
    A, pi = hmm_params.synthetic_hmm(cfg['model'])
    logA_T, log_pi = hmm_params.log_params(A, pi, add_tiny=(cfg['model'] != 'imm'))
    lo, hi = sharding.shard_bounds(cfg['clips'], rank, world)
    n_mine = hi - lo
    algo = a.algo
    if algo == 'dense':
        # the dense kernel for this JOB: for S <= 384 the tensor-memory kernel; for bigger state sets whichever of the
        # tensor-memory and the streaming kernel needs less time for the wave plan HBM allows (passes x clips per pass /
        # measured efficiency of a full pass: 0.607 and 0.747 of the FP32 max-plus peak at S = 722)
        from viterbi_spl_b200.waves import plan_waves, wave_bytes_per_clip
        names = {v: k for k, v in _lib.ALGO_NAMES.items()}
        algo = names.get(_lib.load().vit_select_algo(n_mine, T, S), 'auto')
        if S > 384:
            free, _ = torch.cuda.mem_get_info()
            best = None
            for cand, eff in (('tmem', 0.607), ('stream', 0.747)):
                try:
                    qc = _lib.clips_in_flight(S, _lib.ALGO_NAMES[cand])
                except Exception:
                    continue
                nb = 2 if int(free * 0.85) // wave_bytes_per_clip(T, S, 2) >= qc else 1
                plan = plan_waves(n_mine, wave_bytes_per_clip(T, S, nb), int(free * 0.85), qc)
                cost = sum(-(-(y - x) // qc) for x, y in plan) * qc / eff
                if best is None or cost < best[0]:
                    best = (cost, cand)
            if best:
                algo = best[1]
    dec = ViterbiDecoder(logA_T, log_pi, device=dev, algo=algo)
    wd = WaveDecoder(dec, T, max_wave_clips=a.max_wave_clips)

    checked = {'clips': 0, 'equal': None}
    decode_events = []
    totals = {'frames': 0, 'score_sum': 0.0}

    def fill(start, stop, out):
        synth.device_dense_softmax(stop - start, T, S, seed=10_000 + lo + start, device=dev, out=out)
        return None

    main_stream = torch.cuda.current_stream()

    def sink(start, stop, paths, scores):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(main_stream)
        decode_events[-1].append(ev)
        totals['frames'] += (stop - start) * T
        if start == 0 and a.check > 0 and rank == 0 and 'keep' in checked:
            n = min(a.check, stop - start)       # device-side copies; compared with the oracle after the timed region
            checked['keep'] = (wd._emis[0][:n].clone(), paths[:n].clone(), scores[:n].clone())

    # decode_device is bracketed by events: begin recorded by wrapping the decoder call
    orig = dec.decode_device

    def timed_decode(*args, **kw):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(main_stream)
        decode_events.append([ev])
        return orig(*args, **kw)

    dec.decode_device = timed_decode

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up: one quantum through the same path (module load, workspace growth is part of the first wave otherwise)
    warm = WaveDecoder(dec, T, max_wave_clips=32)
    warm.run(min(n_mine, 32), fill, lambda *x: decode_events[-1].append(None))
    del warm
    decode_events.clear()
    dec._ws = None
    torch.cuda.empty_cache()
    wd = WaveDecoder(dec, T, max_wave_clips=a.max_wave_clips)

    checked['keep'] = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = _lib.launch_count()
    e0.record(main_stream)
    waves = wd.run(n_mine, fill, sink)
    e1.record(main_stream)
    barrier()
    job_ms = e0.elapsed_time(e1)
    keep = checked.pop('keep', None)
    if keep is not None:
        from oracle import np_oracle
        rp, rs = np_oracle.decode_batch_np(logA_T, log_pi, keep[0].cpu().numpy(), None)
        checked['clips'] = int(keep[0].shape[0])
        checked['equal'] = bool(np.array_equal(keep[1].cpu().numpy(), rp) and np.array_equal(keep[2].cpu().numpy(), rs))
    decode_ms = sum(x.elapsed_time(y) for x, y in decode_events)
    t = torch.tensor([job_ms, decode_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    job_ms, decode_ms = float(t[0]), float(t[1])
    frames = cfg['clips'] * T
    if rank == 0:
        cells = S * S if algo in ('tmem', 'stream', 'cluster', 'backpointer') or dec.structure.kind != 1 else None
        line = {
            'metric': 'viterbi_frames_per_sec', 'unit': 'frames/s', 'n_gpus': world, 'scaling': 'strong',
            'config': {'workload': f"config {a.config}: {cfg['clips']} clips x {T} frames x {S} states ({cfg['model']} state set)",
                       'algo': algo, 'structure_kind': int(dec.structure.kind),
                       'halfwidth': int(dec.structure.halfwidth),
                       'emission_bytes_total': cfg['clips'] * T * S * 4,
                       'waves_rank0': [b - a_ for a_, b in waves], 'wave_quantum': wd.quantum,
                       'emission_buffers': wd.emission_buffers,
                       'wave_budget_bytes': wd.budget_bytes},
            'value': frames / (decode_ms * 1e-3), 'decode_ms': decode_ms,
            'job_value_including_generation': frames / (job_ms * 1e-3), 'job_ms': job_ms,
            'frac_of_fp32_maxplus_peak': (frames * cells / (decode_ms * 1e-3) / (world * 148 * 64 * 1.965e9)) if cells else None,
            'gpu_launches': _lib.launch_count() - launches0,
            'parity_vs_oracle': checked, 'data': 'synthetic (generated on the device per wave)',
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
