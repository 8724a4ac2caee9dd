"""Jobs larger than one GPU's HBM, decoded in waves.

BASELINE config 3 (4096 clips x 10,000 frames x 722 states = 118 GB of float32 emissions) and config 5 (65,536 clips x
3000 frames x 361 states = 284 GB) do not fit next to their delta history (the reference's ``T1`` table,
imm/tf_viterbi.py:91), so the job is cut into waves of clips.  Clips are independent (the reference decodes one
recording per call, dcnet/softmax_viterbi.py:3033-3040): a wave is just a smaller batch.  What the planner adds:

* every wave but the last is a multiple of the kernel's *quantum* -- the clips one launch keeps in flight with every SM
  busy (``vit_clips_in_flight``: 1036 for the tensor-memory kernel at S = 361, 1184 for the banded kernels) -- so no
  launch ends with a partly filled pass over the SMs (section 8e: "choose sequences-per-CTA so every wave is full");
* two emission buffers alternate, so wave k+1 is produced (generated on the device, or copied from pinned host memory)
  on a side stream while wave k is decoded -- unless two buffers leave less than one quantum per wave where a single
  buffer holds one (config 3 on the streaming kernel), in which case a full pass beats an overlapped fill.

Host logic only; the compute is ``ViterbiDecoder.decode_device``.
"""
import torch


def plan_waves(n_clips, bytes_per_clip, budget_bytes, quantum=1, max_wave_clips=None):
    """Cut ``n_clips`` into consecutive waves ``[(start, stop), ...]``.

    Each wave holds at most ``budget_bytes // bytes_per_clip`` clips (and at most ``max_wave_clips``), rounded DOWN to a
    multiple of ``quantum`` where that leaves at least one quantum; only the last wave may be ragged."""
    n_clips, quantum = int(n_clips), max(1, int(quantum))
    if n_clips <= 0:
        return []
    cap = int(budget_bytes // max(1, int(bytes_per_clip)))
    if max_wave_clips is not None:
        cap = min(cap, int(max_wave_clips))
    if cap < 1:
        raise MemoryError(f'one clip needs {bytes_per_clip} bytes, the wave budget is {budget_bytes}')
    if cap >= quantum:
        cap -= cap % quantum
    waves, start = [], 0
    while start < n_clips:
        stop = min(n_clips, start + cap)
        waves.append((start, stop))
        start = stop
    return waves


def wave_bytes_per_clip(T, S, emission_buffers=2):
    """Device bytes one clip of a wave occupies: `emission_buffers` emission tables + the delta history + int64 path."""
    return (emission_buffers + 1) * T * S * 4 + T * 8 + 4


class WaveDecoder:
    """Decodes ``n_clips`` clips of ``T`` frames through a ViterbiDecoder, one HBM-sized wave at a time.

        wd = WaveDecoder(decoder, T)
        wd.run(n_clips, fill, sink)

    ``fill(start, stop, out)`` writes the log-emissions of clips ``[start, stop)`` into the CUDA tensor ``out``
    ``[stop - start, T, S]`` on torch's current stream (a side stream here) and returns ``None`` or a CUDA int32
    ``lengths`` tensor; ``sink(start, stop, paths, scores)`` consumes the wave's results (CUDA tensors, valid on the
    current stream; they are reused by the next wave but one).
    """

    def __init__(self, decoder, T, hbm_fraction=0.85, max_wave_clips=None, budget_bytes=None):
        self.dec = decoder
        self.T = int(T)
        S = decoder.S
        from . import _lib
        self.quantum = _lib.clips_in_flight(S, decoder.algo, decoder.structure)
        if budget_bytes is None:
            with torch.cuda.device(decoder.device):
                free, _total = torch.cuda.mem_get_info()
            budget_bytes = int(free * hbm_fraction)
        self.budget_bytes = int(budget_bytes)
        # two emission buffers (wave k+1 is produced while wave k is decoded) unless that leaves less than one quantum
        # per wave where a single buffer would hold one: long clips of a big state set (config 3: 29 MB per clip) on
        # the streaming kernel (2072 clips in flight) -- a full pass beats an overlapped fill
        self.emission_buffers = 2
        if (self.budget_bytes // wave_bytes_per_clip(self.T, S, 2) < self.quantum
                <= self.budget_bytes // wave_bytes_per_clip(self.T, S, 1)):
            self.emission_buffers = 1
        self.bytes_per_clip = wave_bytes_per_clip(self.T, S, self.emission_buffers)
        self.max_wave_clips = max_wave_clips
        self._emis = [None, None]
        self._out = [None, None]
        self._fill_stream = torch.cuda.Stream(device=decoder.device)

    def plan(self, n_clips):
        return plan_waves(n_clips, self.bytes_per_clip, self.budget_bytes, self.quantum, self.max_wave_clips)

    def _buffers(self, slot, n):
        S, T, dev = self.dec.S, self.T, self.dec.device
        if self._emis[slot] is None or self._emis[slot].shape[0] < n:
            self._emis[slot] = None
            self._emis[slot] = torch.empty((n, T, S), dtype=torch.float32, device=dev)
            self._out[slot] = (torch.empty((n, T), dtype=torch.int64, device=dev),
                               torch.empty((n,), dtype=torch.float32, device=dev))
        return self._emis[slot][:n], self._out[slot][0][:n], self._out[slot][1][:n]

    def run(self, n_clips, fill, sink):
        """Returns the wave plan it executed."""
        waves = self.plan(n_clips)
        if not waves:
            return waves
        with torch.cuda.device(self.dec.device):
            main, side = torch.cuda.current_stream(), self._fill_stream
            biggest = max(b - a for a, b in waves)
            nbuf = self.emission_buffers
            for slot in range(min(nbuf, len(waves))):
                self._buffers(slot, biggest)
            side.wait_stream(main)
            decoded = [None, None]           # event: the decode that last read emission buffer `slot` has finished

            def produce(k):
                a, b = waves[k]
                slot = k % nbuf
                emis, _, _ = self._buffers(slot, b - a)
                with torch.cuda.stream(side):
                    if decoded[slot] is not None:
                        side.wait_event(decoded[slot])
                    lengths = fill(a, b, emis)
                    ev = torch.cuda.Event()
                    ev.record(side)
                return lengths, ev

            nxt = produce(0)
            for k, (a, b) in enumerate(waves):
                lengths, ready = nxt
                if nbuf > 1 and k + 1 < len(waves):
                    nxt = produce(k + 1)                 # into the other buffer, while this wave is decoded
                slot = k % nbuf
                emis, paths, scores = self._buffers(slot, b - a)
                main.wait_event(ready)
                self.dec.decode_device(emis, lengths, paths, scores)
                ev = torch.cuda.Event()
                ev.record(main)
                decoded[slot] = ev
                sink(a, b, paths, scores)
                if nbuf == 1 and k + 1 < len(waves):
                    nxt = produce(k + 1)                 # same buffer: ordered after this wave's decode
        return waves
