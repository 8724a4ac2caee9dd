#!/usr/bin/env python
"""Benchmark of the Viterbi hot path (BASELINE.json metric: Viterbi frames/sec, 1024 clips x 3000 frames x 361 states).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--clips B --frames T --states S]

One "step" = one pass of the decoder (forward recursion + backtrace) over one batch of synthetic log-posteriors.
`value`  : whole-job frames/s with the batch already resident in HBM (CUDA events, max over ranks).
`e2e`    : the same through the public host API (ViterbiDecoder.decode_host): pinned host emissions -> device, decode,
           paths + scores -> host, every step.
`roofline`: the forward kernel against the FP32 max-plus issue peak of BASELINE.md section 4 (cells/s), timed with
           CUDA events recorded by the library around that kernel; `roofline_hbm` gives the same launch in GB/s.
`cpu_baseline`: the NumPy restatement of the reference decode (oracle/np_oracle.py, the reference's own CPU path is
           NumPy) on a bounded sample of the same workload over all host cores.
Extra keys of the same line (the other BASELINE.json configurations at their named sizes, each with an oracle check):
`forward_backward` : config 4 -- scaled forward-backward posteriors, 1024 x 3000 x 361 (tcgen05): ms, frames/s, HBM
           fraction, max |gamma - float64 oracle| on 3 clips.
`cfg3_stream`: config 3's kernel -- S = 722 with logA^T streamed from L2 through the bulk-copy (TMA) ring: one full pass
           of 2072 clips x 500 frames, fraction of the FP32 max-plus peak, 2 clips against the oracle.
`strong` : config 5's sweep -- 65,536 clips x 3000 x 361 sharded over the N ranks (65,536 / N clips per GPU, decoded in
           HBM-sized waves generated on the device): total frames/s, i.e. STRONG scaling when run at N = 1, 2, 4, 8.
Under torchrun every rank decodes its own batch of the same shape (clips are independent: weak scaling, no collective
on the data path); rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'viterbi_frames_per_sec'
UNIT = 'frames/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--clips', type=int, default=1024)
    ap.add_argument('--frames', type=int, default=3000)
    ap.add_argument('--states', type=int, default=361)
    ap.add_argument('--algo', default='dense', help="kernel of the headline 'value': dense (= the tensor-memory "
                    "max-plus kernel, what the FP32 roofline is about), or auto/tmem/cluster/backpointer/banded")
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--cpu-seconds', type=float, default=12.0, help='target wall time of the cpu_baseline sample')
    ap.add_argument('--no-extras', action='store_true', help='skip the forward_backward / cfg3_stream / strong legs')
    ap.add_argument('--strong-clips', type=int, default=65536, help='total clips of the strong-scaling leg (config 5)')
    return ap.parse_args()


def shared_config(a):
    """The `config` object -- identical in both arms (`--impl ours` and `--impl reference`): it names the workload; what is
    specific to an arm (kernel, parallelism; the CPU arm's sampling) lives in `details` / `sample`."""
    return {'workload': workload_name(a), 'clips_per_gpu': a.clips, 'frames': a.frames, 'states': a.states,
            'l2': f'no flush needed: one step reads {a.clips * a.frames * a.states * 4 / 1e9:.1f} GB of emissions, far more than the 126 MB L2'}


def workload_name(a):
    return f'batched max-plus Viterbi: {a.clips} clips x {a.frames} frames x {a.states} states (tonet state set), dense transition'


def hmm_for(S):
    from viterbi_spl_b200 import hmm_params
    name = {321: 'dcnet', 361: 'tonet', 722: 'jdc'}.get(S)
    if name is None:
        from viterbi_spl_b200 import synth
        return synth.dyadic_hmm(S, seed=S)
    A, pi = hmm_params.synthetic_hmm(name)
    return hmm_params.log_params(A, pi)


# ---- CPU arm: the reference's NumPy decode, all host cores ------------------------------------------------------

def _cpu_worker(args):
    """One process = one core: decode `n` clips of [T, S] with the NumPy restatement; returns frames decoded."""
    seed, n, T, S = args
    for k in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[k] = '1'
    from oracle import np_oracle
    from viterbi_spl_b200 import synth
    logA_T, log_pi = hmm_for(S)
    frames = 0
    for i in range(n):
        E = synth.dyadic((T, S), seed * 1000 + i)
        np_oracle.viterbi_log_np(logA_T, log_pi, E)
        frames += T
    return frames


def cpu_sample(T, S, clips_per_core, cores):
    """Wall-clock frames/s of `cores` processes each decoding `clips_per_core` clips."""
    import multiprocessing as mp
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(i, 0, T, S) for i in range(cores)])          # import + page-in, untimed
        t0 = time.perf_counter()
        frames = sum(pool.map(_cpu_worker, [(i, clips_per_core, T, S) for i in range(cores)]))
        dt = time.perf_counter() - t0
    return frames / dt, dt, frames


def cpu_baseline(a):
    cores = os.cpu_count() or 1
    # one 3000 x 361 clip costs ~0.45 s of one core; size the sample to ~a.cpu_seconds of wall time
    per_clip_s = 0.45 * (a.frames / 3000.0) * (a.states / 361.0) ** 2
    clips_per_core = max(1, int(a.cpu_seconds / per_clip_s))
    clips_per_core = min(clips_per_core, 24)
    fps, dt, frames = cpu_sample(a.frames, a.states, clips_per_core, cores)
    return {'value': fps, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': f'{cores * clips_per_core} clips x {a.frames} frames x {a.states} states, {clips_per_core} per core, '
                      f'{dt:.1f} s wall; NumPy restatement of imm/tf_viterbi.py:75-109 (TensorFlow not installed; the '
                      f'reference calls its NumPy path the faster one)'}


def run_reference(a):
    """--impl reference: the reference's CPU decode (NumPy port; /root/reference is Python and does not travel)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_clip_s = 0.45 * (a.frames / 3000.0) * (a.states / 361.0) ** 2
    # every step decodes clips_per_core clips on each core; keep the whole run within a few minutes
    budget = 150.0 / max(1, a.steps + a.warmup)
    clips_per_core = max(1, min(8, int(budget / per_clip_s)))
    import multiprocessing as mp
    ctx = mp.get_context('fork')
    times = []
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(i, 0, a.frames, a.states) for i in range(cores)])
        for step in range(a.warmup + a.steps):
            t0 = time.perf_counter()
            frames = sum(pool.map(_cpu_worker, [(step * cores + i, clips_per_core, a.frames, a.states) for i in range(cores)]))
            dt = time.perf_counter() - t0
            if step >= a.warmup:
                times.append(dt)
    total = sum(times)
    fps = frames * len(times) / total
    sample = (f'{cores * clips_per_core} clips x {a.frames} frames x {a.states} states per step '
              f'({clips_per_core} per core x {cores} processes, all host cores whatever --gpus is); NumPy restatement of '
              f'imm/tf_viterbi.py:75-109 (the reference is pure Python/NumPy; TensorFlow is not installed)')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
        'warmup': a.warmup, 'ms_per_step': 1e3 * total / len(times), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': shared_config(a), 'sample': sample,
        'cpu_baseline': {'value': fps, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': fps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---- clocks sampler ---------------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), f'--query-gpu={self.FIELDS}',
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.12)
        self.proc.terminate()     # exact PID we started
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        inside = [l for (ts, l) in self.lines if t_begin <= ts <= t_end] or [l for (_, l) in self.lines]
        for l in inside:
            parts = [p.strip() for p in l.split(',')]
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except Exception:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': sorted(reasons)}


# ---- GPU arm ----------------------------------------------------------------------------------------------------

def bind_to_gpu_numa_node(local_rank):
    """Pin this rank to the CPUs of its GPU's NUMA node BEFORE any page-locked staging buffer is allocated, so that the
    4.4 GB of host emissions of every rank sit in memory local to that GPU's PCIe root (first-touch placement).  With 8
    ranks uploading at once, remote-node buffers halve the aggregate host->device bandwidth.  Best effort: silently does
    nothing where sysfs does not say."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        dev = f'{getattr(pr, "pci_domain_id", 0):04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0'
        with open(f'/sys/bus/pci/devices/{dev}/local_cpulist') as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(','):
            if '-' in part:
                a, b = part.split('-')
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None



# ---- the other BASELINE.json configurations (extra keys of the line) -----------------------------------------------

def leg_forward_backward(dev, world, barrier, all_max, peaks, steps=3):
    """Config 4: scaled forward-backward posteriors, 1024 clips x 3000 frames x 361 states per GPU (tcgen05 kernel)."""
    import torch
    from viterbi_spl_b200 import ForwardBackward, _lib, hmm_params
    B, T, S = 1024, 3000, 361
    A, pi = hmm_params.synthetic_hmm('tonet')
    A, pi = A.astype(np.float32), pi.astype(np.float32)
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    lik = torch.softmax(2.0 * torch.randn((B, T, S), device=dev, generator=g), dim=-1)    # dense softmax likelihoods
    gamma = torch.empty_like(lik)
    ll = torch.empty(B, device=dev)

    def timed(impl):
        fb = ForwardBackward(A, pi, device=dev, impl=impl)
        fb.run_device(lik, None, gamma, ll)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        n0 = _lib.launch_count()
        e0.record()
        for _ in range(steps):
            fb.run_device(lik, None, gamma, ll)
        e1.record()
        barrier()
        return all_max(e0.elapsed_time(e1) / steps), _lib.launch_count() - n0

    def parity():
        try:
            from oracle import fb_oracle
            sub = [0, B // 2, B - 1]
            wg, wl = fb_oracle.forward_backward_batch_np(A, pi, lik[sub].cpu().numpy())
            return {'max_abs_gamma_err': float(np.abs(gamma[sub].cpu().numpy() - wg).max()),
                    'max_rel_loglik_err': float(np.abs((ll[sub].cpu().numpy() - wl) / wl).max()),
                    'clips_checked': len(sub), 'tolerance': 'gamma 1e-4 abs, log L 1e-5 rel'}
        except Exception as ex:
            return f'oracle unavailable: {ex}'

    hbm = float(peaks.get('hbm_gbs', 6650.0))
    bytes_ = 20.0 * B * T * S           # b read twice, alpha~ written and read, gamma written
    # the structured fast path first (what impl='auto' takes for this matrix: band +-14 built from one jump histogram +
    # the unvoiced state, exact zeros elsewhere -- S (2d+3) instead of S^2 multiply-adds per frame, plain fp32 FFMA),
    # then the headline: the dense tcgen05 kernel config 4 names, whose results stay in gamma / ll for the parity check
    ms_b, launches_b = timed('banded')
    structured = {'impl': 'banded (scaled-Toeplitz band as a convolution, one warp per clip: csrc/vit_fb_conv.cu)', 'ms_per_step': ms_b, 'value': world * B * T / (ms_b * 1e-3), 'unit': UNIT,
                  'gpu_launches': int(launches_b), 'halfwidth': 14, 'dtype': 'f32 (FFMA)',
                  'roofline_hbm': {'bound': 'hbm', 'achieved': bytes_ / (ms_b * 1e-3) / 1e9, 'peak': hbm, 'unit': 'GB/s',
                                   'frac': bytes_ / (ms_b * 1e-3) / 1e9 / hbm, 'algorithmic_bytes': bytes_},
                  'parity_vs_float64_oracle': parity()}
    ms, launches = timed('tc')
    out = {'workload': f'scaled forward-backward posteriors: {B} clips x {T} frames x {S} states per GPU', 'dtype': 'f32 '
           '(products as 2 x bf16 terms on tcgen05, fp32 accumulate)', 'ms_per_step': ms, 'steps': steps,
           'value': world * B * T / (ms * 1e-3), 'unit': UNIT, 'gpu_launches': int(launches)}
    out['roofline_hbm'] = {'bound': 'hbm', 'achieved': bytes_ / (ms * 1e-3) / 1e9, 'peak': hbm, 'unit': 'GB/s',
                           'frac': bytes_ / (ms * 1e-3) / 1e9 / hbm, 'algorithmic_bytes': bytes_}
    flops = 2.0 * 2.0 * B * (T - 1) * S * S       # fwd + bwd, 2 flop per cell (useful fp32-equivalent work)
    out['useful_tflops'] = flops / (ms * 1e-3) / 1e12
    out['parity_vs_float64_oracle'] = parity()
    out['structured_fast_path'] = structured
    return out


def leg_cfg3_stream(dev, world, barrier, all_max, peaks, steps=3):
    """Config 3's kernel: S = 722, dense (jdc matrix run through the DENSE recursion), logA^T streamed from L2 through
    the 4-stage bulk-copy (TMA) ring.  One full pass of the streaming kernel -- 2072 clips = 148 SMs x 14 -- x 500 frames
    (the named 4096 x 10,000 job is this launch repeated over waves and frame ranges: tools/bench_waves.py --config 3)."""
    import torch
    from viterbi_spl_b200 import ViterbiDecoder, _lib, hmm_params, synth
    B, T, S = 2072, 500, 722
    A, pi = hmm_params.synthetic_hmm('jdc')
    logA_T, log_pi = hmm_params.log_params(A, pi)
    dec = ViterbiDecoder(logA_T, log_pi, device=dev, algo='stream')
    emis = synth.device_dense_softmax(B, T, S, seed=33, device=dev)
    host = synth.batch('dense_softmax', 2, T, S, seed0=500)
    emis[:2] = torch.as_tensor(host).to(dev)
    paths = torch.empty((B, T), dtype=torch.int64, device=dev)
    scores = torch.empty((B,), dtype=torch.float32, device=dev)
    dec.decode_device(emis, None, paths, scores)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    n0 = _lib.launch_count()
    e0.record()
    for k in range(steps):
        dec.decode_device(emis, None, paths, scores, forward_events=ev[k])
    e1.record()
    barrier()
    launches = _lib.launch_count() - n0
    ms = all_max(e0.elapsed_time(e1) / steps)
    fwd = all_max(statistics.mean(x.elapsed_time(y) for x, y in ev))
    mhz = float(peaks.get('sm_max_mhz', 1965.0))
    cells = float(B) * (T - 1) * S * S
    out = {'workload': f'dense max-plus Viterbi, {B} clips x {T} frames x {S} states per GPU (jdc state set), logA^T '
                       'streamed from L2 via cp.async.bulk (TMA) ring', 'algo': 'stream', 'ms_per_step': ms,
           'forward_ms': fwd, 'steps': steps, 'value': world * B * T / (ms * 1e-3), 'unit': UNIT,
           'gpu_launches': int(launches),
           'roofline': {'bound': 'fp32_alu', 'kernel': 'stream_forward_kernel', 'achieved': cells / (fwd * 1e-3) / 1e12,
                        'peak': 148 * 64 * mhz * 1e6 / 1e12, 'unit': 'Tcell/s', 'frac': cells / (fwd * 1e-3) / (148 * 64 * mhz * 1e6)}}
    try:
        from oracle import c_oracle
        rp, rs = c_oracle.decode_batch_c(logA_T, log_pi, host)
        out['parity_vs_oracle'] = bool(np.array_equal(rp, paths[:2].cpu().numpy()) and np.array_equal(rs, scores[:2].cpu().numpy()))
    except Exception as ex:
        out['parity_vs_oracle'] = f'oracle unavailable: {ex}'
    return out


def leg_strong(a, dev, rank, world, barrier, all_max, peaks):
    """Config 5's sweep: `--strong-clips` (65,536) clips x 3000 frames x 361 states in total, sharded contiguously over
    the ranks (65,536 / N per GPU), each shard decoded in HBM-sized waves whose emissions are generated on the device
    on a side stream while the previous wave is decoded.  The total work is fixed, so the value at N = 1, 2, 4, 8 is the
    STRONG-scaling curve.  `value` counts the decode calls alone (CUDA events, max over ranks); `job_value` the whole
    job including whatever generation did not hide."""
    import torch
    from viterbi_spl_b200 import ViterbiDecoder, _lib, hmm_params, sharding, synth
    from viterbi_spl_b200.waves import WaveDecoder
    T, S = 3000, 361
    A, pi = hmm_params.synthetic_hmm('tonet')
    logA_T, log_pi = hmm_params.log_params(A, pi)
    lo, hi = sharding.shard_bounds(a.strong_clips, rank, world)
    n_mine = hi - lo
    dec = ViterbiDecoder(logA_T, log_pi, device=dev, algo='tmem')
    main = torch.cuda.current_stream()
    begin, end, keep = [], [], {}
    orig = dec.decode_device

    def timed_decode(*args, **kw):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(main)
        begin.append(ev)
        return orig(*args, **kw)

    dec.decode_device = timed_decode

    def fill(start, stop, out):
        synth.device_dense_softmax(stop - start, T, S, seed=10_000 + lo + start, device=dev, out=out)
        return None

    def sink(start, stop, paths, scores):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(main)
        end.append(ev)
        if start == 0 and rank == 0 and 'wd' in keep:
            keep['clips'] = (keep['wd']._emis[0][:2].clone(), paths[:2].clone(), scores[:2].clone())

    warm = WaveDecoder(dec, T, max_wave_clips=32)               # module load, first workspace
    warm.run(min(n_mine, 32), fill, sink)
    del warm
    begin.clear(), end.clear()
    dec._ws = None
    torch.cuda.empty_cache()
    wd = WaveDecoder(dec, T)
    keep['wd'] = wd
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    n0 = _lib.launch_count()
    e0.record(main)
    waves = wd.run(n_mine, fill, sink)
    e1.record(main)
    barrier()
    launches = _lib.launch_count() - n0
    job_ms = all_max(e0.elapsed_time(e1))
    decode_ms = all_max(sum(x.elapsed_time(y) for x, y in zip(begin, end)))
    frames = float(a.strong_clips) * T
    mhz = float(peaks.get('sm_max_mhz', 1965.0))
    out = {'workload': f'{a.strong_clips} clips x {T} frames x {S} states in total, {n_mine} per GPU, dense max-plus (tmem kernel), '
                       'waves generated on the device', 'scaling': 'strong', 'n_gpus': world,
           'value': frames / (decode_ms * 1e-3), 'unit': UNIT, 'decode_ms': decode_ms, 'job_ms': job_ms,
           'job_value': frames / (job_ms * 1e-3), 'waves_rank0': [y - x for x, y in waves], 'wave_quantum': wd.quantum,
           'gpu_launches': int(launches),
           'frac_of_fp32_maxplus_peak': frames * (T - 1) / T * S * S / (decode_ms * 1e-3) / (world * 148 * 64 * mhz * 1e6)}
    if rank == 0 and 'clips' in keep:
        try:
            from oracle import c_oracle
            E2, p2, s2 = keep['clips']
            rp, rs = c_oracle.decode_batch_c(logA_T, log_pi, E2.cpu().numpy())
            out['parity_vs_oracle'] = bool(np.array_equal(rp, p2.cpu().numpy()) and np.array_equal(rs, s2.cpu().numpy()))
        except Exception as ex:
            out['parity_vs_oracle'] = f'oracle unavailable: {ex}'
    keep.clear()
    return out


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the ncu capture under profiles/ (tools/ncu_summary.py traffic), but only if
    that capture was taken from THIS build of the kernel (hash of its sources, viterbi_spl_b200.build.kernel_build_id)."""
    from viterbi_spl_b200 import build
    try:
        with open(os.path.join(ROOT, 'profiles', 'forward_traffic.json')) as fh:
            rec = json.load(fh).get('kernels', {}).get(kernel)
    except Exception:
        rec = None
    if rec is None:
        return None, 'no ncu capture of this kernel under profiles/forward_traffic.json'
    now = build.kernel_build_id(kernel)
    if rec.get('build_id') != now:
        return None, f"stale ncu capture (taken from build {rec.get('build_id')}, this is {now}): not reported"
    return rec, None


def run_ours(a):
    import torch
    import torch.distributed as dist
    from viterbi_spl_b200 import ViterbiDecoder, _lib, synth

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    # stdout carries exactly ONE JSON line: anything a library prints there meanwhile (NCCL's version banner) goes
    # to stderr instead
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    # CPU baseline first (N=1 only): worker processes are forked before this process creates a CUDA context
    cpu = cpu_baseline(a) if (world == 1 and not a.no_cpu) else None
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl ours needs a CUDA device (there is no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    numa_cpus = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def all_max(x):
        t_ = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_[0])

    B, T, S = a.clips, a.frames, a.states
    logA_T, log_pi = hmm_for(S)
    # the headline runs the DENSE max-plus recursion (S^2 cells per frame -- the work BASELINE.json's roofline counts);
    # the bit-exact structured fast path that algo='auto' would take for this banded matrix is reported separately
    algo = a.algo
    if algo == 'dense':
        # (vit_select_algo names the DENSE kernel auto would take for this shape: tmem, or stream for big state sets)
        algo = {v: k for k, v in _lib.ALGO_NAMES.items()}.get(_lib.load().vit_select_algo(a.clips, a.frames, a.states), 'auto')
    dec = ViterbiDecoder(logA_T, log_pi, device=dev, algo=algo)
    emis = synth.device_dense_softmax(B, T, S, seed=1234 + rank, device=dev)      # 4.4 GB: far larger than the 126 MB L2
    paths = torch.empty((B, T), dtype=torch.int64, device=dev)
    scores = torch.empty((B,), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream()

    # ---- device-resident throughput ("value") + forward-kernel time ("roofline") -----------------------------
    # (PipelinedDecoder -- backtrace of step k under the forward of step k+1 -- was measured and buys nothing here: the
    # forward kernel keeps the dispatch ports 92 % busy, so the overlapped backtrace just slows it by its own cost)
    for _ in range(a.warmup):
        dec.decode_device(emis, None, paths, scores)
    fwd_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.2)
    barrier()
    launches0 = _lib.launch_count()
    t_begin = time.perf_counter()
    e0.record(stream)
    for k in range(a.steps):
        dec.decode_device(emis, None, paths, scores, forward_events=fwd_ev[k])
    e1.record(stream)
    barrier()
    t_end = time.perf_counter()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    total_ms = e0.elapsed_time(e1)
    fwd_ms = statistics.mean(x.elapsed_time(y) for (x, y) in fwd_ev)
    t = torch.tensor([total_ms, fwd_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, fwd_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / a.steps
    frames_per_rank = B * T
    value = world * frames_per_rank / (ms_per_step * 1e-3)

    # ---- the structured fast path on the same inputs (same results bit for bit, S (2d+2) cells per frame) -----------
    structured = None
    if dec.structure.kind == 1 and algo != 'banded':
        dec_b = ViterbiDecoder(logA_T, log_pi, device=dev, algo='banded')
        pb = torch.empty_like(paths)
        sb = torch.empty_like(scores)
        for _ in range(a.warmup):
            dec_b.decode_device(emis, None, pb, sb)
        ev_b = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        b0.record(stream)
        for k in range(a.steps):
            dec_b.decode_device(emis, None, pb, sb, forward_events=ev_b[k])
        b1.record(stream)
        barrier()
        tb = torch.tensor([b0.elapsed_time(b1) / a.steps, statistics.mean(x.elapsed_time(y) for (x, y) in ev_b)],
                          dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        structured = {'algo': 'banded', 'value': world * frames_per_rank / (float(tb[0]) * 1e-3), 'unit': UNIT,
                      'ms_per_step': float(tb[0]), 'forward_ms': float(tb[1]),
                      'halfwidth': int(dec.structure.halfwidth), 'dense_state': int(dec.structure.dense_index),
                      'identical_to_dense': bool(torch.equal(pb, paths) and torch.equal(sb, scores)),
                      'forward_hbm_GBps': float(B) * T * S * 8 / (float(tb[1]) * 1e-3) / 1e9}
        del dec_b, pb, sb

    # ---- end-to-end through the host API ------------------------------------------------------------------------
    e2e = None
    if not a.no_e2e:
        host = torch.empty((B, T, S), dtype=torch.float32).pin_memory()
        host.copy_(emis)
        e2e_steps = max(1, min(a.steps, 5))
        out_p = torch.empty((B, T), dtype=torch.int64).pin_memory()       # the caller's page-locked result buffers
        out_s = torch.empty((B,), dtype=torch.float32).pin_memory()
        for _ in range(1):
            dec.decode_host(host, out=(out_p, out_s))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            hp, hs = dec.decode_host(host, out=(out_p, out_s))
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {'value': world * frames_per_rank * e2e_steps / float(tt[0]), 'unit': UNIT,
               'h2d_bytes_per_step': int(world * B * T * S * 4), 'd2h_bytes_per_step': int(world * (B * T * 8 + B * 4)),
               'steps': e2e_steps, 'api': 'ViterbiDecoder.decode_host (pinned host emissions in, NumPy paths+scores out in page-locked memory)'}
        del host

    # ---- parity spot check of the timed configuration against the oracle (not timed) ------------------------------
    parity = None
    if rank == 0:
        try:
            from oracle import c_oracle
            nchk = min(B, 16)
            ref_p, ref_s = c_oracle.decode_batch_c(logA_T, log_pi, emis[:nchk].cpu().numpy())
            parity = bool(np.array_equal(ref_p, paths[:nchk].cpu().numpy()) and
                          np.array_equal(ref_s, scores[:nchk].cpu().numpy()))
        except Exception as ex:   # the oracle is test infrastructure; its absence must not break the bench
            parity = f'oracle unavailable: {ex}'

    # ---- the other configurations of BASELINE.json at their named sizes (every rank takes part) -----------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    extras = {}
    algo_res = dec.algo if dec.algo else _lib.load().vit_select_algo(B, T, S)
    structure_kind = int(dec.structure.kind)
    if not a.no_extras:
        del dec, emis, paths, scores
        torch.cuda.empty_cache()
        for name, fn in (('forward_backward', lambda: leg_forward_backward(dev, world, barrier, all_max, peaks)),
                         ('cfg3_stream', lambda: leg_cfg3_stream(dev, world, barrier, all_max, peaks)),
                         ('strong', lambda: leg_strong(a, dev, rank, world, barrier, all_max, peaks))):
            try:
                extras[name] = fn()
            except Exception as ex:           # an extra leg must not take the headline down with it
                if world > 1:
                    raise
                extras[name] = {'error': f'{type(ex).__name__}: {ex}'}
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline (forward kernel) ------------------------------------------------------------------------------
    sm_mhz_peak = float(peaks.get('sm_max_mhz', 1965.0))
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured' if peaks else 'fallback'
    cells = float(B) * (T - 1) * S * S                      # cells = sum_b (T_b - 1) S^2, 1 add + 1 max each
    cells_per_s = cells / (fwd_ms * 1e-3)
    alu_peak = 148 * 64 * sm_mhz_peak * 1e6                 # N_SM x 128 lanes x f / 2 instr per cell
    # algorithmic HBM bytes per launch: SURVEY 8(d) counts 6 S per state-frame (emissions 4 B in + uint16 backpointers
    # out); this kernel writes the fp32 delta history instead (4 B: index tracking in the hot loop costs 2.7x the cells/clk,
    # profiles/r01_microbench_pipes.jsonl) = 8 S -- both stated, `traffic_ratio` is against SURVEY's definition
    algo_bytes_survey = float(B) * T * S * 6
    algo_bytes = float(B) * T * S * 8
    algo_name = {1: 'backpointer', 2: 'cluster', 3: 'tmem', 4: 'banded', 5: 'stream'}.get(algo_res, str(algo_res))
    kernel_name = {1: 'bp_forward_kernel', 2: 'cluster_forward_kernel', 3: 'tmem_forward_kernel',
                   4: 'banded_forward_kernel', 5: 'stream_forward_kernel'}.get(algo_res, '?')
    rec, traffic_note = measured_traffic(kernel_name)
    traffic = None
    if rec is not None:
        # the capture's launch shape may differ in clips/frames from the timed one: traffic is linear in state-frames
        traffic = float(rec['dram_bytes_per_launch']) * (float(B) * T * S) / (float(rec['clips']) * rec['frames'] * rec['states'])
    roofline = {'bound': 'fp32_alu', 'kernel': kernel_name, 'achieved': cells_per_s / 1e12,
                'peak': alu_peak / 1e12, 'unit': 'Tcell/s', 'frac': cells_per_s / alu_peak,
                'peak_definition': f'148 SMs x 64 cells/clk (FADD+FMNMX, 2 issue slots per cell) x {sm_mhz_peak:.0f} MHz '
                                   f'({peak_src} sm_max_mhz)',
                'kernel_ms': fwd_ms, 'kernel_share_of_step': fwd_ms / ms_per_step, 'traffic': traffic,
                'algorithmic_bytes': {'survey_6S_uint16_backpointers': algo_bytes_survey, 'kernel_8S_fp32_delta_history': algo_bytes},
                'traffic_ratio': (traffic / algo_bytes_survey) if traffic else None,
                'traffic_source': ({'ncu_capture': rec.get('source'), 'build_id': rec.get('build_id'),
                                    'capture_shape': [rec['clips'], rec['frames'], rec['states']]} if rec else traffic_note)}
    roofline_hbm = {'bound': 'hbm', 'achieved': algo_bytes / (fwd_ms * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                    'frac': algo_bytes / (fwd_ms * 1e-3) / 1e9 / hbm_peak, 'traffic': traffic,
                    'note': f'{peak_src} copy bandwidth; the kernel is FP32-issue bound, not HBM bound'}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': shared_config(a),
        'details': {'algo': algo_name, 'parallelism': f'{world} x independent clip shards, no data-path collective',
                    'host_numa_binding': numa_cpus},
        'roofline': roofline, 'roofline_hbm': roofline_hbm,
        'structured_fast_path': structured, 'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks, 'parity_vs_oracle': parity,
    }
    if cpu is not None:
        line['cpu_baseline'] = cpu
    line.update(extras)
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)


if __name__ == '__main__':
    main()
