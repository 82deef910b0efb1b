"""One long recording decoded as overlapping segments, one per GPU (BASELINE
configs[2]: "20-min 48 kHz recording split into line-aligned overlapping segments across 2/4/8 GPUs").

The reference's ``process()`` (wefax.py:46-93) couples the whole recording in three places:

1. ``scipy.signal.resample`` / ``hilbert`` are ONE circular transform of the recording
   (wefax.py:384,174).  Here every segment is transformed on its own with a halo on both sides that is
   thrown away afterwards: the only approximation of this mode (tolerance: DESIGN.md section 6).
2. ``numpy.percentile(env, (0.5, 99.5))`` are global order statistics (wefax.py:196).  Kept EXACT
   over the union of the segment cores: three rounds of 4 x 2048-bin radix-digit histograms
   (``wefax_segment_histogram``) summed over all segments on the host.
3. ``start_frame`` (wefax.py:80) shifts every image row.  Kept EXACT given the grey levels: the
   phasing search reads only the first <= 100 peaks, i.e. the head of segment 0
   (``wefax_segment_sync``), and its result is broadcast.  Image line ``r`` then belongs to the
   segment whose core holds its first sample ``start_frame + r*w``: that is the line alignment; the
   halo (>= 3 lines) supplies the two neighbouring lines above / below the x4 bicubic needs.

No collective on the data path and no NCCL: ranks exchange 32 KiB histograms, two doubles and one
integer through the host (``torch.distributed`` object / tensor collectives on whatever backend the
process group has), then rank 0 gathers the image rows.

The protocol is written against a small worker interface (``segment_envelope / segment_histogram /
segment_quantise / segment_sync / segment_raster``: ``Decoder`` implements it on the GPU); a process
drives one worker per local GPU context.
"""
from __future__ import annotations

import math
import os
import sys
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Callable, Sequence

import numpy as np

from . import _native as N

DEFAULT_HALO = 65536          # 11025-Hz samples discarded on each inner side of a segment
MARGIN_LINES = 2              # bicubic support of the x4 vertical resize (Pillow: 2 input lines)


@dataclass(frozen=True)
class Segment:
    """All positions are in the recording's own coordinates; the halo of the first / last segment wraps
    around the recording's ends (negative ``in_begin`` / ``in_end`` beyond the recording), because the
    reference's transforms are circular."""
    index: int
    in_begin: int       # input frames [in_begin, in_end) of the extended segment
    in_end: int
    out_begin: int      # 11025-Hz sample where the extended segment starts
    n_out: int          # 11025-Hz samples of the extended segment
    core_begin: int     # 11025-Hz samples this segment owns
    core_end: int
    seam: int = 0       # position inside the extended segment where the recording's end meets its start

    @property
    def out_end(self) -> int:
        return self.out_begin + self.n_out


def segment_frames(pcm, seg: Segment, n_frames: int | None = None):
    """The input frames of the extended segment (numpy array or torch tensor), wrapped where needed."""
    n = int(pcm.shape[0]) if n_frames is None else n_frames
    if seg.in_begin >= 0 and seg.in_end <= n:
        return pcm[seg.in_begin:seg.in_end]
    if seg.in_begin < 0:
        parts = [pcm[n + seg.in_begin:], pcm[:seg.in_end]]
    else:
        parts = [pcm[seg.in_begin:], pcm[:seg.in_end - n]]
    if type(pcm).__module__.startswith("torch"):
        import torch
        return torch.cat(parts)
    return np.concatenate(parts)


def _smooth(n: int) -> bool:
    for p in (2, 3, 5, 7, 11, 13):
        while n % p == 0:
            n //= p
    return n == 1


def _plannable(n: int) -> bool:
    """True when the real-input transform of length n runs on the direct (non-Bluestein) plan."""
    if n < 2 or n % 2 or not _smooth(n):
        return False
    try:
        return N.fft_plan_describe(n // 2)[1] == 0
    except ValueError:
        return False


def plan_segments(n_frames: int, sample_rate: int, n_segments: int, halo: int = DEFAULT_HALO,
                  plannable: Callable[[int], bool] | None = _plannable, max_grow: int = 16384, head: int = 0) -> list:
    """Cut ``n_frames`` input frames into ``n_segments`` cores with a halo of >= ``halo`` 11025-Hz samples
    on both sides (circular at the recording's ends).  Cut points are multiples of
    ``sample_rate / gcd(sample_rate, 11025)`` input frames so that they fall on whole 11025-Hz samples
    (48 kHz: 640 frames <-> 147 samples).  Halos are grown (up to ``max_grow`` units) until the extended
    lengths factor into the transform's radices.  ``head``: 11025-Hz samples the FIRST extended segment
    must reach (the phasing search's horizon)."""
    if n_segments < 1:
        raise ValueError("n_segments must be >= 1")
    sample_rate = int(sample_rate)
    n_total = n_frames if sample_rate == N.TARGET_RATE else N.resampled_length(n_frames, sample_rate)
    if n_segments == 1:
        return [Segment(0, 0, n_frames, 0, n_total, 0, n_total)]
    g = math.gcd(sample_rate, N.TARGET_RATE)
    u_in, u_out = sample_rate // g, N.TARGET_RATE // g
    units = n_frames // u_in
    if n_frames % u_in or units * u_out != n_total:
        raise ValueError(f"segment mode needs a recording of a whole number of {u_in}-frame units that resamples "
                         f"to exactly {u_out} samples each (got {n_frames} frames -> {n_total} samples)")
    halo_u = -(-int(halo) // u_out)
    halo16_u = -(-halo_u // 16) * 16          # keeps the first segment's own samples 64-byte aligned
    if units < n_segments or units // n_segments + 2 * halo16_u >= units:
        raise ValueError("recording too short for that many segments with this halo")
    cuts = [round(j * units / n_segments) for j in range(n_segments + 1)]

    def ext_ok(b, e):
        if plannable is None:
            return True
        return plannable((e - b) * u_out) and (sample_rate == N.TARGET_RATE or plannable((e - b) * u_in))

    segs = []
    for j in range(n_segments):
        c0, c1 = cuts[j], cuts[j + 1]
        first, last = j == 0, j == n_segments - 1
        b = -halo16_u if first else c0 - halo_u
        e = units + halo_u if last else c1 + halo_u
        if first:
            e = max(e, -(-int(head) // u_out))

        # grow the free sides until the transform lengths are plannable; failing that the library falls back
        # to its chirp-z transform
        def candidates(b=b, e=e, first=first, last=last):
            yield b, e
            for grow in range(1, max_grow + 1):
                if not last:
                    yield b, e + grow
                if not first:
                    yield b - grow, e
        b, e = next((c for c in candidates() if c[1] - c[0] < units and ext_ok(*c)), (b, e))
        if e - b >= units:                   # the extended segment would lap itself: take the whole recording
            b, e = 0, units
        seam = -b * u_out if b < 0 else ((units - b) * u_out if e > units else 0)
        segs.append(Segment(j, b * u_in, e * u_in, b * u_out, (e - b) * u_out, c0 * u_out, c1 * u_out, seam))
    return segs


def search_horizon(lpm) -> int:
    """How far (11025-Hz samples) the first extended segment must reach so that the phasing search finds its
    100 peaks inside it: 101 line periods of at most frame + 500 samples (wefax.py:251,264-266)."""
    lc = N.line_constants(float(lpm))
    return (N.MAX_PEAKS + 1) * int(math.ceil(lc["dev_max"])) + lc["template_len"]


def plan_decode(n_frames: int, sample_rate: int, lpm, n_segments: int, halo: int = DEFAULT_HALO,
                plannable: Callable[[int], bool] | None = _plannable, head: int | None = None) -> list:
    """``plan_segments`` with the head the phasing search of this LPM needs (what ``decode_segmented`` uses)."""
    return plan_segments(n_frames, sample_rate, n_segments, halo, plannable,
                         head=search_horizon(lpm) if head is None else head)


# --------------------------------------------------------------------------------------------- #
# exact global percentiles from per-segment digit histograms                                       #
# --------------------------------------------------------------------------------------------- #
def percentile_targets(n_total: int):
    """numpy 'linear' (wefax.py:196): virtual index ``(N-1)*(q/100)``; returns the four order-statistic
    ranks ``[lo(0.5), lo+1, lo(99.5), lo+1]`` and the two interpolation weights."""
    ranks, fracs = [], []
    for q in (0.5, 99.5):
        virt = (n_total - 1) * (q / 100)
        lo = int(math.floor(virt))
        ranks += [lo, min(lo + 1, n_total - 1)]
        fracs.append(virt - lo)
    return ranks, fracs


def _lerp(a: float, b: float, t: float) -> float:
    d = b - a
    return b - d * (1 - t) if t >= 0.5 else a + d * t


def select_order_statistics(histogram: Callable[[int, Sequence[int]], np.ndarray], ranks: Sequence[int]) -> list:
    """``histogram(level, prefix) -> (4, 2048)`` counts ALREADY SUMMED over all segments.  Narrows each
    of the four ranks digit by digit (11 + 11 + 10 bits of the float32 bit pattern) and returns the
    four order statistics as float32 values."""
    rem = [int(r) for r in ranks]
    prefix = [0, 0, 0, 0]
    for level, bits in ((0, 11), (1, 11), (2, 10)):
        h = np.asarray(histogram(level, prefix), dtype=np.int64)
        for t in range(4):
            row = h[0] if level == 0 else h[t]
            cum = np.cumsum(row)
            b = int(np.searchsorted(cum, rem[t], side="right"))
            if b >= (1 << bits) or b >= row.shape[0]:
                raise RuntimeError("percentile rank beyond the histogram population "
                                   "(NaN or negative envelope values?)")
            rem[t] -= int(cum[b - 1]) if b else 0
            prefix[t] = (prefix[t] << bits) | b
    return [float(np.array([p], dtype=np.uint32).view(np.float32)[0]) for p in prefix]


# --------------------------------------------------------------------------------------------- #
# host exchange                                                                                    #
# --------------------------------------------------------------------------------------------- #
class HostExchange:
    """Sum / broadcast / gather between the ranks of a ``torch.distributed`` group (or nothing when
    there is a single process).  Payloads are tiny; tensors go to the GPU only if the backend is NCCL."""

    def __init__(self, group=None, device: int | None = None):
        self.group, self.device = group, device
        # a process group can only exist if torch.distributed is already imported: never pay for the import here
        dist = sys.modules.get("torch.distributed")
        self.dist = dist if dist is not None and dist.is_available() and dist.is_initialized() else None
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.rank = self.dist.get_rank(group) if self.dist else 0

    def device_exchange_possible(self, local_workers: int) -> bool:
        """The device-resident percentile exchange needs one worker per rank and a NCCL group (WEFAX_SEG_HOST=1
        forces the host exchange, for A/B measurements)."""
        return (self.world > 1 and local_workers == 1 and self.dist.get_backend(self.group) == "nccl"
                and os.environ.get("WEFAX_SEG_HOST", "0") != "1")

    def device_state(self, device: int):
        import torch
        key = ("state", device)
        if getattr(self, "_state_key", None) != key:
            self._state = torch.zeros(N.SEG_STATE_WORDS, dtype=torch.int32, device=f"cuda:{device}")
            self._state_key = key
        return self._state

    def sum(self, arr: np.ndarray) -> np.ndarray:
        if self.world == 1:
            return arr
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64))
        if self.dist.get_backend(self.group) == "nccl":
            t = t.cuda(self.device)
        self.dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def broadcast(self, obj, src: int = 0):
        if self.world == 1:
            return obj
        box = [obj if self.rank == src else None]
        self.dist.broadcast_object_list(box, src=src, group=self.group)
        return box[0]

    def broadcast_ints(self, values, count: int, src: int = 0) -> list:
        """One fixed-size int64 tensor broadcast (cheaper than pickling an object through two collectives)."""
        if self.world == 1:
            return [int(v) for v in values]
        import torch
        t = torch.zeros(count, dtype=torch.int64)
        if self.rank == src:
            t[:len(values)] = torch.tensor([int(v) for v in values], dtype=torch.int64)
        nccl = self.dist.get_backend(self.group) == "nccl"
        if nccl:
            t = t.cuda(self.device)
        self.dist.broadcast(t, src=src, group=self.group)
        return t.cpu().tolist()

    def gather_rows_on_device(self, rows: dict, spans: Sequence, width: int, n_rows: int, local_workers: int):
        """Image rows that live in each GPU's memory -> one image on rank 0's GPU (NCCL send / recv over NVLink),
        then a single copy into pinned host memory.  spans: ``(first_row, end_row)`` of every segment, known to
        every rank from the plan, so no metadata travels.  Returns the host image on rank 0, None elsewhere."""
        import torch
        if self.world > 1 and self.dist.get_backend(self.group) != "nccl":
            raise RuntimeError("rows_on_device needs a NCCL process group (gloo has no GPU send / recv); "
                               "gather host rows instead (rows_on_device=False)")
        if self.device is None:
            raise ValueError("rows_on_device needs HostExchange(device=<this rank's GPU>)")
        img = None
        if self.rank == 0:
            key = (n_rows, width)
            if getattr(self, "_img_key", None) != key:
                self._img = torch.empty(key, dtype=torch.uint8, device=f"cuda:{self.device}")
                self._img_host = torch.empty(key, dtype=torch.uint8, pin_memory=True)
                self._img_key = key
            img = self._img
        for index, (y0, y1) in enumerate(spans):
            if y1 <= y0:
                continue
            owner = index // local_workers
            if self.rank == 0 and owner == 0:
                img[y0:y1].copy_(rows[y0], non_blocking=True)
            elif self.rank == 0:
                self.dist.recv(img[y0:y1], src=owner, group=self.group)
            elif owner == self.rank:
                self.dist.send(rows[y0], dst=0, group=self.group)
        if self.rank != 0:
            return None
        self._img_host.copy_(img)
        torch.cuda.current_stream().synchronize()
        return self._img_host.numpy()

    def gather(self, obj, dst: int = 0):
        if self.world == 1:
            return [obj]
        bucket = [None] * self.world if self.rank == dst else None
        self.dist.gather_object(obj, bucket, dst=dst, group=self.group)
        return bucket


# --------------------------------------------------------------------------------------------- #
# the protocol                                                                                     #
# --------------------------------------------------------------------------------------------- #
@dataclass
class SegmentedResult:
    n_out: int
    width: int
    low: float
    high: float
    peaks: list
    phasing_signals: list
    start_frame: int
    status: int
    rows: dict                       # {first image row (x4): (rows, width) uint8} of the local segments
    digitalized: dict                # {core_begin: uint8 array} of the local segments (if asked for)
    demodulated: dict
    image: np.ndarray | None = None  # the assembled (4h, w) raster on the gathering rank

    def error(self):
        if self.status & N.REC_NAN:
            return ValueError("cannot convert float NaN to integer")       # wefax.py:216
        if self.status & N.REC_NO_GROUPS:
            return ValueError("max() iterable argument is empty")           # wefax.py:294
        if self.status & N.REC_NO_LINES:
            return IndexError("image index out of range")                   # wefax.py:304
        return None


def owned_lines(seg: Segment, start_frame: int, width: int, n_lines_total: int):
    """Image lines whose first sample ``start_frame + r*width`` lies in the segment's core."""
    def first_line_at_or_after(pos):
        return max(0, -(-(pos - start_frame) // width))
    r0 = min(first_line_at_or_after(seg.core_begin), n_lines_total)
    r1 = min(first_line_at_or_after(seg.core_end), n_lines_total)
    return r0, r1


def decode_segmented(pcm, sample_rate: int, lpm, workers: Sequence, n_segments: int | None = None,
                     halo: int = DEFAULT_HALO, notch_freq=2600, notch_q=1, want=("raster",),
                     exchange: HostExchange | None = None, gather: bool = True,
                     plannable: Callable[[int], bool] | None = _plannable,
                     segment_pcm: Callable[[Segment], object] | None = None,
                     n_frames: int | None = None, head: int | None = None,
                     rows_on_device: bool = False, segments: Sequence[Segment] | None = None,
                     parallel: bool = True) -> SegmentedResult:
    """Decode ONE recording in ``world * len(workers)`` segments; this process drives ``workers`` (one
    per local GPU context) on segments ``rank*len(workers) ...``.

    pcm: the whole recording (int16 ``(n,)`` or ``(n, 2)``), or None with ``segment_pcm(seg)`` returning what
    ``segment_frames(pcm, seg)`` would (host array or CUDA tensor) and ``n_frames`` the total length.
    want: any of ``raster``, ``digitalized``, ``demodulated``.  With ``gather`` the image rows are
    collected on rank 0 (``result.image``).  head: how far the first extended segment must reach for the
    phasing search (default: 101 line periods + 500 samples each, wefax.py:251,264-266).
    segments: a plan made by ``plan_decode`` with the same arguments (e.g. to stage the PCM beforehand).
    rows_on_device: leave each worker's image rows in its GPU's memory (torch uint8 tensors); with ``gather``
    they travel GPU to GPU (NCCL send / recv) to rank 0 and reach the host in one pinned copy.
    parallel: drive several local workers from one host thread each (a single process over all GPUs of a host
    needs no process group at all)."""
    ex = exchange or HostExchange()
    L = len(workers)
    G = n_segments or ex.world * L
    if G != ex.world * L:
        raise ValueError(f"{G} segments need {G} workers over all ranks (have {ex.world} x {L})")
    if n_frames is None:
        n_frames = int(pcm.shape[0])
    sample_rate = int(sample_rate)
    n_total = n_frames if sample_rate == N.TARGET_RATE else N.resampled_length(n_frames, sample_rate)
    lc = N.line_constants(float(lpm))
    w = lc["width"]
    segs = list(segments) if segments is not None else plan_decode(n_frames, sample_rate, lpm, G, halo, plannable, head)
    if len(segs) != G:
        raise ValueError(f"plan has {len(segs)} segments, expected {G}")
    mine = segs[ex.rank * L:(ex.rank + 1) * L]
    if G > 1 and halo < (MARGIN_LINES + 1) * w:
        raise ValueError(f"halo {halo} shorter than {MARGIN_LINES + 1} lines of {w} samples")

    # Several local workers (one per GPU of a single process, or several contexts on one GPU) run every phase
    # concurrently, one host thread each: the native calls release the GIL and each context is only ever used by
    # one thread at a time.
    pool = ThreadPoolExecutor(L) if (parallel and L > 1) else None

    def each(fn):
        pairs = list(zip(workers, mine))
        return list(pool.map(lambda p: fn(*p), pairs)) if pool else [fn(*p) for p in pairs]

    try:
        return _run_protocol(ex, each, workers, segs, L, pcm, sample_rate, lpm, w, n_total, n_frames, segment_pcm,
                             notch_freq, notch_q, want, gather, rows_on_device)
    finally:
        if pool:
            pool.shutdown(wait=True)


def _run_protocol(ex, each, workers, segs, L, pcm, sample_rate, lpm, w, n_total, n_frames, segment_pcm, notch_freq,
                  notch_q, want, gather, rows_on_device) -> SegmentedResult:
    # 1. envelopes of the extended segments (resident on each GPU)
    def envelope(wk, sg):
        part = segment_pcm(sg) if segment_pcm is not None else segment_frames(pcm, sg, n_frames)
        n_ext = wk.segment_envelope(part, sample_rate, sg.core_begin - sg.out_begin, sg.core_end - sg.out_begin,
                                    notch_freq, notch_q, n_out=sg.n_out, seam=sg.seam)
        if n_ext != sg.n_out:
            raise RuntimeError(f"segment {sg.index}: {n_ext} samples at 11025 Hz, planned {sg.n_out}")
    each(envelope)

    # 2. exact global percentiles
    ranks, fracs = percentile_targets(n_total)
    digitalized, demodulated = {}, {}
    small = tuple(x for x in want if x in ("digitalized", "demodulated"))
    if ex.device_exchange_possible(L):
        # Everything stays on the devices: per radix level a histogram kernel, a 32 KiB all-reduce of the counters
        # (NCCL, queued on the context's stream: control data, not the compute path) and a select kernel that every
        # rank runs redundantly; the grey map reads low / high from device memory.  No host round trip, no sync.
        import torch
        wk, sg = workers[0], segs[ex.rank * L]
        with torch.cuda.device(wk.device), torch.cuda.stream(torch.cuda.ExternalStream(wk.stream, device=wk.device)):
            state = ex.device_state(wk.device)
            wk.segment_select_init(state, ranks, fracs[0], fracs[1])
            for level in (0, 1, 2):
                wk.segment_histogram_dev(level, state)
                ex.dist.all_reduce(state[N.SEG_STATE_HIST:], group=ex.group)
                wk.segment_select_dev(level, state)
            got = wk.segment_quantise_dev(state, small)
            head = state[:N.SEG_STATE_HIST].cpu()          # low / high / status for the result (a local copy, no exchange)
        low, high = (float(x) for x in head[12:16].numpy().view(np.float64))
        status = int(head[16]) & N.REC_NAN
        if "digitalized" in got:
            digitalized[sg.core_begin] = got["digitalized"]
        if "demodulated" in got:
            demodulated[sg.core_begin] = got["demodulated"]
    else:
        # through the host: three histogram exchanges (any backend, several local workers, or no process group)
        def summed_histogram(level, prefix):
            # one extra counter travels with the histogram: ranks whose kernels failed.  Every rank then raises
            # together instead of leaving the others blocked in the collective.
            local = np.zeros(4 * 2048 + 1, dtype=np.int64)
            failure = None
            try:
                for h in each(lambda wk, sg: wk.segment_histogram(level, prefix)):
                    local[:-1] += h.reshape(-1)
            except Exception as exc:   # noqa: BLE001 - reported below, on every rank
                failure = exc
                local[:] = 0
                local[-1] = 1
            total = ex.sum(local)
            if total[-1]:
                raise RuntimeError(f"segment histogram failed on {int(total[-1])} rank(s)") from failure
            return total[:-1].reshape(4, 2048)

        v = select_order_statistics(summed_histogram, ranks)
        low, high = _lerp(v[0], v[1], fracs[0]), _lerp(v[2], v[3], fracs[1])
        status = N.REC_OK if high != low else N.REC_NAN      # wefax.py:216: 0/0 -> nan -> int() raises

        # 3. grey map everywhere; phasing search on the segment that starts the recording
        for sg, got in each(lambda wk, sg: (sg, wk.segment_quantise(low, high, small))):
            if "digitalized" in got:
                digitalized[sg.core_begin] = got["digitalized"]
            if "demodulated" in got:
                demodulated[sg.core_begin] = got["demodulated"]
    sync = None
    too_short = "first segment too short for the phasing search (it found fewer than 100 peaks before its end); " \
                "pass a larger head"
    # one int64 tensor: [status, start_frame, n_peaks, n_phasing, peaks..., phasing...]; status -1 / -2 = rank 0
    # could not produce it (every rank raises, nobody is left waiting in the broadcast)
    flat = []
    if ex.rank == 0:
        try:
            sync = workers[0].segment_sync(lpm)
            if len(sync["peaks"]) < N.MAX_PEAKS and segs[0].out_end < n_total:
                flat = [-1, 0, 0, 0]   # the search ran off the end of segment 0 before its 100th peak
            else:
                flat = [sync["status"], sync["start_frame"], len(sync["peaks"]), len(sync["phasing_signals"]),
                        *sync["peaks"], *sync["phasing_signals"]]
        except Exception:   # noqa: BLE001
            flat = [-2, 0, 0, 0]
            if ex.world == 1:
                raise
    flat = ex.broadcast_ints(flat, 4 + 2 * N.MAX_PEAKS, 0)
    if flat[0] == -1:
        raise ValueError(too_short)
    if flat[0] == -2:
        raise RuntimeError("the phasing search failed on rank 0")
    sync = {"status": flat[0], "start_frame": flat[1], "peaks": flat[4:4 + flat[2]],
            "phasing_signals": flat[4 + flat[2]:4 + flat[2] + flat[3]]}
    status |= sync["status"]
    start = sync["start_frame"]

    # 4. image lines by ownership of their first sample, with the bicubic margin taken from the halo
    n_lines = (n_total - start) // w                      # wefax.py:299
    rows = {}

    def raster(wk, sg):
        r0, r1 = owned_lines(sg, start, w, n_lines)
        if r1 <= r0:
            return None
        top = min(MARGIN_LINES, r0)
        bottom = min(MARGIN_LINES, n_lines - r1)
        first = start + (r0 - top) * w - sg.out_begin
        out = None
        if rows_on_device:
            import torch
            out = torch.empty((4 * (r1 - r0), w), dtype=torch.uint8, device=f"cuda:{wk.device}")
        return 4 * r0, wk.segment_raster(lpm, first, top + (r1 - r0) + bottom, top, r1 - r0, out=out)

    if "raster" in want and status == N.REC_OK:
        rows = dict(r for r in each(raster) if r is not None)
    if n_lines == 0:
        status |= N.REC_NO_LINES
    res = SegmentedResult(n_total, w, low, high, sync["peaks"], sync["phasing_signals"], start, status, rows,
                          digitalized, demodulated)
    if gather and "raster" in want and rows_on_device and status == N.REC_OK:
        spans = [tuple(4 * r for r in owned_lines(sg, start, w, n_lines)) for sg in segs]
        res.image = ex.gather_rows_on_device(rows, spans, w, 4 * n_lines, L)
    elif gather and "raster" in want:
        parts = ex.gather(rows, 0)
        if ex.rank == 0 and status == N.REC_OK:
            img = np.empty((4 * n_lines, w), dtype=np.uint8)
            filled = 0
            for part in parts:
                for y0, block in part.items():
                    img[y0:y0 + block.shape[0]] = block
                    filled += block.shape[0]
            if filled != 4 * n_lines:
                raise RuntimeError(f"segments delivered {filled} of {4 * n_lines} image rows")
            res.image = img
    return res
