"""Batch sharding of one global batch across the GPUs of a box (BASELINE config 4).

Utterances are independent in eval mode, so the data path needs exactly two collectives: rank 0 packs the inputs of the
global batch into one row per utterance ON THE GPU and scatters equal row blocks (``scatter``); every rank runs
``ZeroVox.forward`` on its block of consecutive utterances; every rank packs the VALID part of its results — the
``mel_len[i] * hop`` samples and ``mel_len[i]`` frames per utterance that the consumer keeps
(utils/export_hifigan.py:138-151) — into one contiguous buffer that rank 0 receives directly at its final offset in the
gathered buffer (``gather-v``: one grouped ncclSend / ncclRecv, no staging copy, no re-ordering: blocks are consecutive).
``torch.distributed`` is the plumbing (NCCL over NVLink / NVSwitch on GPUs; the same code runs over gloo with CPU tensors,
which is how the host logic is tested — the ragged pack / unpack then use plain tensor slicing instead of the CUDA
kernels of csrc/ragged.cu).

Host-to-host delivery.  When the batch comes from and the waveforms go back to HOST memory, funnelling everything through
rank 0's GPU makes rank 0's one PCIe link carry world x the bytes (8 GPUs, configs[1]: 202 MB of waveforms per step = 4 ms,
a fifth of the step).  :class:`SharedHostBuffer` is a page-locked host window mapped by every rank of the group;
with ``host_out=SharedHostBuffer`` every rank writes its own waveforms to the host over its OWN PCIe link (only the small
mel / log-duration tails still travel to rank 0 over NCCL, which is also the completion fence), and with
``x=SharedHostBatch`` every rank uploads its own block of the global batch from the window (no scatter).

Control plane: one broadcast of 12 int64 (shapes, flags, the global frame count when durations are forced) unless every
rank passes ``spec``; with predicted durations one all-gather of the per-utterance frame counts, which also yields the
global frame count.

Exactness against an unsharded run.  The reference's batch-composition quirks make an utterance's tail
depend on the batch's padded lengths: every shard keeps the *global* phoneme length T (inputs are never trimmed), decodes
and vocodes at the *global* frame count ``L_pad`` (``pad_to``), and takes the reference's "zero-fill padded mel frames iff
a mel mask exists and B > 1" decision (model.py:283-285) from the GLOBAL batch size (``zero_padded_mel``).  The valid
samples / frames of every utterance are then those of the unsharded batch.  ``tails="valid"`` (default) returns zeros past
an utterance's own length; ``tails="padded"`` also ships the padded tails, reproducing the unsharded tensors everywhere.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import tempfile
from dataclasses import dataclass, field
from typing import Callable, Mapping, Sequence

import torch
import torch.distributed as dist

_HEADER = 12                                           # int64 words
_ALIGN = 64                                            # elements: every rank's segment of the gathered buffer starts 256-byte aligned


def partition(n_utterances: int, world: int) -> list[range]:
    """Blocks of consecutive utterances, ceil(B / world) per rank (the last ranks may get fewer or none).  Every shard is
    padded to the global T and L_pad, so a shard's cost is proportional to its utterance COUNT whatever the utterance
    lengths are: equal counts are balanced, and consecutive blocks let rank 0 receive every result in place."""
    nb = -(-n_utterances // world) if n_utterances else 0
    return [range(min(r * nb, n_utterances), min((r + 1) * nb, n_utterances)) for r in range(world)]


# ---------------------------------------------------------------------------------------------------------------------
# ragged pack / unpack: CUDA kernels through the C ABI on the GPU, tensor slicing for the gloo / CPU plumbing tests
# ---------------------------------------------------------------------------------------------------------------------
def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _ragged_copy(pack: bool, padded: torch.Tensor, packed: torch.Tensor, lens_dev: torch.Tensor, offs_dev: torch.Tensor,
                 lens: Sequence[int], offs: Sequence[int], rows: int, unit: int, zero_tail: bool = True):
    """padded [B, rows, max_units * unit] (or [B, max_units * unit] when rows == 1) <-> packed 1-D."""
    B = padded.shape[0]
    if B == 0:
        return
    p3 = padded.reshape(B, rows, -1)
    max_units = p3.shape[2] // unit
    if padded.is_cuda:
        from . import _lib
        lib = _lib.load()
        assert padded.is_contiguous() and packed.is_contiguous() and padded.dtype == packed.dtype == torch.float32
        st = C.c_void_p(torch.cuda.current_stream(padded.device).cuda_stream)
        with torch.cuda.device(padded.device):
            if pack:
                rc = lib.zvx_ragged_pack(_ptr(p3), p3.stride(0), p3.stride(1), rows, unit, B, max_units, _ptr(lens_dev),
                                         _ptr(offs_dev), _ptr(packed), st)
            else:
                rc = lib.zvx_ragged_unpack(_ptr(packed), _ptr(lens_dev), _ptr(offs_dev), rows, unit, B, max_units, _ptr(p3),
                                           p3.stride(0), p3.stride(1), 1 if zero_tail else 0, st)
        if rc != 0:
            raise RuntimeError("zvx_ragged: " + lib.zvx_ragged_last_error().decode())
        return
    for b in range(B):                                 # gloo / CPU plumbing path (tests)
        n = max(0, min(int(lens[b]), max_units)) * unit
        seg = packed[offs[b]: offs[b] + rows * n].view(rows, n)
        if pack:
            seg.copy_(p3[b, :, :n])
        else:
            p3[b, :, :n] = seg
            if zero_tail:
                p3[b, :, n:] = 0


# ---------------------------------------------------------------------------------------------------------------------
# host memory shared by the ranks of a box
# ---------------------------------------------------------------------------------------------------------------------
def _group_info(group):
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if not multi:
        return False, 0, 1, 0
    return True, dist.get_rank(group), dist.get_world_size(group), (0 if group is None else dist.get_global_rank(group, 0))


class SharedHostBuffer:
    """``nbytes`` of host memory mapped by every rank of ``group`` (one process per GPU on ONE box) and page-locked for DMA
    in each of them, so that every GPU reads / writes it over its own PCIe link.  Collective: every rank of the group
    constructs it in the same order.  Backed by an unlinked tmpfs file (/dev/shm, else the temp directory); with CUDA the
    mapping is registered with ``cudaHostRegister`` (non-blocking copies then go straight to / from it); without
    (gloo tests) it is plain shared memory.

    ``tensor(dtype, numel, offset_bytes)`` returns a view.  Ordering is the caller's: a rank may read what another wrote
    only after something that orders the two (sharded_forward's tail exchange does that for the waveforms)."""

    def __init__(self, nbytes: int, group=None, register: bool | None = None):
        multi, rank, world, root = _group_info(group)
        self.nbytes = nbytes = max(int(nbytes), 16)
        self.group, self.rank, self.world = group, rank, world
        path = None
        if rank == 0:
            d = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > nbytes + (64 << 20) else None
            fd, path = tempfile.mkstemp(prefix="zvx_host_", dir=d)
            os.ftruncate(fd, nbytes)
            os.close(fd)
        if multi:
            box = [path]
            dist.broadcast_object_list(box, src=root, group=group)
            path = box[0]
        try:
            self.bytes = torch.from_file(path, shared=True, size=nbytes, dtype=torch.uint8)
        finally:
            if multi:
                dist.barrier(group=group)              # every rank has mapped it
            if rank == 0:
                os.unlink(path)
        self.registered = False
        if register is None:
            register = torch.cuda.is_available()
        if register:
            rc = torch.cuda.cudart().cudaHostRegister(self.bytes.data_ptr(), nbytes, 0)
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister of the shared host window failed: {rc}")
            self.registered = True

    def tensor(self, dtype=torch.float32, numel: int | None = None, offset_bytes: int = 0) -> torch.Tensor:
        item = torch.empty((), dtype=dtype).element_size()
        if numel is None:
            numel = (self.nbytes - offset_bytes) // item
        if offset_bytes % item or offset_bytes + numel * item > self.nbytes:
            raise ValueError("view outside the shared host window")
        return self.bytes[offset_bytes: offset_bytes + numel * item].view(dtype)

    def close(self):
        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self.bytes.data_ptr())
            self.registered = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SharedHostBatch:
    """The global input batch in a :class:`SharedHostBuffer`: rank 0 passes the batch dict (host tensors), the other
    ranks ``None``; every rank then holds ``.x`` — the same dict as views of the shared window — and
    ``sharded_forward(model, batch)`` lets every rank upload its own block of utterances (no NCCL scatter, no 8x traffic on
    rank 0's PCIe link).  Rank 0 refills the window in place for the next batch (``.x[name].copy_(...)``; same shapes) and
    publishes it with whatever tells the other ranks that a batch is ready (the header broadcast does when ``spec`` is not
    used)."""

    def __init__(self, x: Mapping[str, torch.Tensor] | None, group=None, register: bool | None = None):
        multi, rank, world, root = _group_info(group)
        meta = None
        if rank == 0:
            meta, off = [], 0
            for k, v in x.items():
                if isinstance(v, torch.Tensor):
                    meta.append((k, v.dtype, tuple(v.shape), off))
                    off = -(-(off + v.numel() * v.element_size()) // 256) * 256
            meta = (meta, off)
        if multi:
            box = [meta]
            dist.broadcast_object_list(box, src=root, group=group)
            meta = box[0]
        fields, total = meta
        self.window = SharedHostBuffer(total, group=group, register=register)
        self.x = {}
        for k, dt, shape, off in fields:
            n = 1
            for d in shape:
                n *= d
            self.x[k] = self.window.tensor(dt, n, off).view(shape)
            if rank == 0:
                self.x[k].copy_(x[k])
        if multi:
            dist.barrier(group=group)                  # filled before anyone reads

    def nbytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in self.x.values())


# ---------------------------------------------------------------------------------------------------------------------
# the gathered result on rank 0
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class RaggedBatch:
    """Rank 0's gathered results: ``buf`` holds, per rank in utterance order, [wav | mel | log_duration] of that rank's
    block — the valid ``mel_len[i] * hop`` samples and ``[n_mels, mel_len[i]]`` frames of every utterance."""
    buf: torch.Tensor                                  # float32 1-D
    B: int
    T: int
    L: int                                             # global padded frame count
    hop: int
    n_mels: int
    lens: list[int]                                    # frames shipped per utterance (mel_len, or L with tails="padded")
    mel_len_host: list[int]
    wav_off: list[int]                                 # element offsets into buf, per utterance
    mel_off: list[int]
    logd_off: list[int]
    seg: list[tuple[int, int, int]]                    # per rank: (segment start, wav elements, segment elements)
    gather_bytes: int = 0                              # bytes received from the other ranks
    scatter_bytes: int = 0                             # bytes sent to the other ranks
    events: dict = field(default_factory=dict)
    host: torch.Tensor | None = None                   # host_out delivery: the host buffer (fp32, 1-D) ...
    host_wav_off: list[int] | None = None              # ... and every utterance's offset in it
    wav_on_device: bool = True                         # False: SharedHostBuffer delivery — peers' waveforms exist on the host only

    def wav(self, i: int) -> torch.Tensor:
        if not self.wav_on_device:
            raise RuntimeError("the waveforms were delivered to the shared host buffer only: use host_wav(i)")
        return self.buf[self.wav_off[i]: self.wav_off[i] + self.lens[i] * self.hop]

    def host_wav(self, i: int) -> torch.Tensor:
        """Utterance i's valid samples in the host buffer (after the caller has synchronised the stream)."""
        if self.host is None:
            raise RuntimeError("no host_out was given")
        return self.host[self.host_wav_off[i]: self.host_wav_off[i] + self.lens[i] * self.hop]

    def mel(self, i: int) -> torch.Tensor:
        return self.buf[self.mel_off[i]: self.mel_off[i] + self.lens[i] * self.n_mels].view(self.n_mels, self.lens[i])

    def log_duration(self) -> torch.Tensor:
        return torch.stack([self.buf[o: o + self.T] for o in self.logd_off]) if self.B else self.buf.new_zeros((0, self.T))

    def wav_segments(self) -> list[tuple[int, int]]:
        """(start, end) element ranges of buf that hold waveforms — one per rank (what an e2e caller copies to the host)."""
        return [(s, s + w) for s, w, _ in self.seg if w]

    def padded(self):
        """The tuple of ``ZeroVox.forward`` for the whole batch: (wav [B, L*hop], mel [B, n_mels, L], mel_len int64 [B],
        log_duration [B, T]); zeros past every utterance's shipped length."""
        dev = self.buf.device
        if not self.wav_on_device:
            raise RuntimeError("the waveforms were delivered to the shared host buffer only: use host_wav(i)")
        wav = torch.empty((self.B, self.L * self.hop), dtype=torch.float32, device=dev)
        mel = torch.empty((self.B, self.n_mels, self.L), dtype=torch.float32, device=dev)
        lens_d = torch.tensor(self.lens, dtype=torch.int64, device=dev)
        _ragged_copy(False, wav, self.buf, lens_d, torch.tensor(self.wav_off, dtype=torch.int64, device=dev), self.lens,
                     self.wav_off, 1, self.hop)
        _ragged_copy(False, mel, self.buf, lens_d, torch.tensor(self.mel_off, dtype=torch.int64, device=dev), self.lens,
                     self.mel_off, self.n_mels, 1)
        return wav, mel, torch.tensor(self.mel_len_host, dtype=torch.int64, device=dev), self.log_duration()


def _layout(counts: Sequence[int], lens: Sequence[int], T: int, hop: int, n_mels: int):
    """Offsets of every utterance / rank inside the gathered buffer.  lens: frames shipped per utterance, global order."""
    wav_off, mel_off, logd_off, seg = [], [], [], []
    base, i = 0, 0
    for n in counts:
        F = sum(lens[i:i + n])
        w0, m0, d0 = base, base + F * hop, base + F * (hop + n_mels)
        acc = 0
        for j in range(n):
            wav_off.append(w0 + acc * hop)
            mel_off.append(m0 + acc * n_mels)
            logd_off.append(d0 + j * T)
            acc += lens[i + j]
        size = -(-(F * (hop + n_mels) + n * T) // _ALIGN) * _ALIGN
        seg.append((base, F * hop, size))
        base += size
        i += n
    return wav_off, mel_off, logd_off, seg, base


# ---------------------------------------------------------------------------------------------------------------------
# scatter -> forward -> gather-v
# ---------------------------------------------------------------------------------------------------------------------
def _row_layout(T, T_ref, n_mels_in, has_mask, has_dur, Lm):
    """Byte offsets of the fields of one utterance's input row (4-byte fields first, byte masks last, 16-byte rows)."""
    o, f = 0, {}
    for name, n in (("phoneme", 4 * T), ("puncts", 4 * T), ("duration", 4 * T if has_dur else 0),
                    ("ref_mel", 4 * T_ref * n_mels_in), ("phoneme_mask", T if has_mask else 0), ("mel_mask", Lm)):
        f[name] = (o, o + n)
        o += n
    return f, -(-o // 16) * 16


def sharded_forward(model: Callable, x: Mapping[str, torch.Tensor] | None, force_duration: bool = False,
                    group=None, device: torch.device | str | None = None, hop_length: int = 256, n_mels: int = 80,
                    tails: str = "valid", ragged: bool = False, spec: Sequence[int] | None = None,
                    events: dict | None = None, vocoder_groups: int = 1,
                    host_out: "torch.Tensor | SharedHostBuffer | None" = None):
    """Run ``model(x_shard, force_duration=..., pad_to=..., zero_padded_mel=...)`` on every rank of ``group`` for the
    global batch ``x`` held by rank 0 (other ranks pass ``x=None``).  ``model`` is a ``ZeroVox`` (or any callable with
    that signature returning ``(wav [n, L*hop], mel [n, n_mels, L], mel_len int64 [n], log_duration [n, T])``).

    Returns on rank 0 the tuple of ``ZeroVox.forward`` for the whole batch in the original utterance order, padded to
    the global L_pad (``ragged=True``: the :class:`RaggedBatch` itself, no expansion); ``None`` on the other ranks.
    ``spec``: the 12-word header when every rank already knows it (skips the broadcast); ``events``: a dict that
    receives CUDA events at the phase boundaries (bench.py).

    Pipelined delivery: with ``vocoder_groups`` = G > 1 every rank vocodes its block in G utterance groups and ships each
    group's waveforms as soon as they are enqueued (the gather-v becomes G grouped ncclSend/ncclRecv exchanges on a side
    stream, overlapping the next group's kernels); with ``host_out`` (rank 0: a pinned fp32 buffer) the gathered waveforms
    are also copied to the host group by group — rank-major, utterance order, valid samples back to back — so that only the
    last group's transfer is exposed.

    Host-to-host without the rank-0 funnel: ``host_out`` = a :class:`SharedHostBuffer` given by EVERY rank makes each rank
    copy its own waveforms to the host itself (fixed slots of ``ceil(B / world) * L * hop`` samples per rank, valid samples
    of the rank's utterances back to back; ``RaggedBatch.host_wav(i)``); the waveforms of the other ranks then never reach
    rank 0's GPU (``RaggedBatch.wav_on_device`` is False).  ``x`` = a :class:`SharedHostBatch` given by every rank replaces
    the scatter by each rank uploading its own block from the shared window.
    """
    if tails not in ("valid", "padded"):
        raise ValueError("tails must be 'valid' or 'padded'")
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if multi else (0, 1)
    if device is not None:
        dev = torch.device(device)
    elif multi and dist.get_backend(group) == "nccl" or (not multi and torch.cuda.is_available()):
        dev = torch.device("cuda", torch.cuda.current_device())
    else:
        dev = torch.device("cpu")

    def mark(name):
        if events is not None and dev.type == "cuda":
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            events[name] = e

    shared_in = isinstance(x, SharedHostBatch)
    shared_out = isinstance(host_out, SharedHostBuffer)
    if shared_in:
        x = x.x
    host_t = host_out.tensor(torch.float32) if shared_out else host_out

    # ---- header -------------------------------------------------------------------------------------------------
    all_lens = None                                    # rank 0, forced durations: frames of every utterance
    if rank == 0:
        assert x is not None, "rank 0 must hold the global batch"
        B, T = x["phoneme"].shape
        T_ref, n_mels_in = x["ref_mel"].shape[1], x["ref_mel"].shape[2]
        has_mask, has_dur = "phoneme_mask" in x, bool(force_duration) and "duration" in x
        has_mm = has_dur and "mel_mask" in x
        L_hint, Lm = -1, 0
        if has_dur:
            all_lens = x["duration"].clamp(min=0).sum(1).tolist()
            L_hint = int(max(all_lens)) if all_lens else 0
        if has_mm:
            Lm = x["mel_mask"].shape[1]
        # model.py:283-285 on the GLOBAL batch: zero-fill iff a mel mask exists and B > 1
        zero_pad = int(((not has_dur) or has_mm) and B > 1)
        hdr_list = [B, T, T_ref, n_mels_in, int(has_mask), int(has_dur), int(has_mm), L_hint, Lm, zero_pad,
                    int(tails == "padded"), 0]
        if spec is not None and list(spec) != hdr_list:
            raise ValueError(f"spec {list(spec)} does not describe the batch {hdr_list}")
    if spec is not None:
        hdr_list = [int(v) for v in spec]
    elif multi:
        hdr = torch.tensor(hdr_list if rank == 0 else [0] * _HEADER, dtype=torch.int64).to(dev)
        dist.broadcast(hdr, src=0 if group is None else dist.get_global_rank(group, 0), group=group)
        hdr_list = hdr.cpu().tolist()
    B, T, T_ref, n_mels_in, has_mask, has_dur, has_mm, L_hint, Lm, zero_pad, tails_padded, _ = hdr_list
    has_mask, has_dur, has_mm = bool(has_mask), bool(has_dur), bool(has_mm)
    parts = partition(B, world)
    counts = [len(p) for p in parts]
    nb, n_mine = max(counts) if counts else 0, counts[rank]

    # ---- ONE scatter of the inputs, packed on the device as one row per utterance --------------------------------------
    fields, row_bytes = _row_layout(T, T_ref, n_mels_in, has_mask, has_dur, Lm if has_mm else 0)
    scatter_bytes = 0
    if multi and shared_in:
        # every rank uploads its own block from the shared host window (the model's forward does the H2D, as at N = 1)
        blk = parts[rank]
        xs = {k: v[blk.start: blk.stop] for k, v in x.items() if isinstance(v, torch.Tensor)}
        if not has_dur:
            xs.pop("duration", None)
        mark("packed")
    elif multi:
        recv = torch.empty((nb, row_bytes), dtype=torch.uint8, device=dev)
        send = None
        if rank == 0:
            rows = torch.zeros((world * nb, row_bytes), dtype=torch.uint8, device=dev)
            for name, dt in (("phoneme", torch.int32), ("puncts", torch.int32), ("duration", torch.int32),
                             ("ref_mel", torch.float32), ("phoneme_mask", torch.uint8), ("mel_mask", torch.uint8)):
                a, b = fields[name]
                if b > a:
                    t = x[name].to(dev, non_blocking=True).to(dt).contiguous()
                    rows[:B, a:b] = t.view(torch.uint8).reshape(B, b - a)
            send = list(rows.view(world, nb * row_bytes).unbind(0))
            scatter_bytes = (world - 1) * nb * row_bytes
        mark("packed")
        dist.scatter(recv.view(-1), send, src=0 if group is None else dist.get_global_rank(group, 0), group=group)
        xs = {}
        for name, dt in (("phoneme", torch.int32), ("puncts", torch.int32), ("duration", torch.int32),
                         ("ref_mel", torch.float32), ("phoneme_mask", torch.uint8), ("mel_mask", torch.uint8)):
            a, b = fields[name]
            if b > a:
                t = recv[:n_mine, a:b].contiguous().view(dt)
                xs[name] = t.reshape(n_mine, T_ref, n_mels_in) if name == "ref_mel" else (
                    t.to(torch.bool) if dt == torch.uint8 else t)
    else:
        xs = {k: v for k, v in x.items() if isinstance(v, torch.Tensor)}
    mark("scattered")

    # ---- forward on the block, at the global frame count; results shipped group by group --------------------------------
    from .tts.model import group_bounds
    G = int(vocoder_groups or 1)   # > 1: equal groups; < -1: halving groups (tts.model.group_bounds)
    if G == 0 or G == -1:
        G = 1
    cuda = dev.type == "cuda"
    side = torch.cuda.Stream(dev) if cuda and (G != 1 or host_t is not None) else None
    root = 0 if (group is None or not multi) else dist.get_global_rank(group, 0)
    lens_box, st = {}, {"groups_done": 0}

    def pad_to(local_lmax: int, mel_len_host: Sequence[int]) -> int:
        lens_box["mine"] = [int(v) for v in mel_len_host]
        if has_dur or not multi:
            return max(L_hint, local_lmax)
        return _exchange_lengths(lens_box, nb, world, dev, group)

    def prepare(L, mel_len):
        """Everything that needs the frame counts: shipped lengths, offsets, rank 0's result buffer (lazily, at the first
        finished group: predicted durations are only known after the encoder)."""
        my_true = lens_box.get("mine", [])
        if n_mine > 0 and len(my_true) != n_mine:      # a model that never called pad_to (not ZeroVox): read the lengths
            my_true = [int(v) for v in mel_len.cpu().tolist()]
        st["my_ship"] = [L] * n_mine if tails_padded else my_true
        st["F"] = F = sum(st["my_ship"])
        st["my_size"] = -(-(F * (hop_length + n_mels) + n_mine * T) // _ALIGN) * _ALIGN
        st["L"] = L
        slot = nb * L * hop_length                     # shared host window: a rank's fixed slot (L is global)
        st["host_base"] = rank * slot
        if shared_out and host_t.numel() < world * slot:
            raise ValueError(f"the shared host window holds {host_t.numel()} samples, {world} slots of {slot} are needed")
        if rank == 0:
            if has_dur:
                true_all = [int(v) for v in all_lens]
            elif multi:
                true_all = [int(v) for v in lens_box["all"]]
            else:
                true_all = list(my_true)
            ship_all = [L] * B if tails_padded else true_all
            wav_off, mel_off, logd_off, seg, total = _layout(counts, ship_all, T, hop_length, n_mels)
            buf = torch.empty(total, dtype=torch.float32, device=dev)
            st["out"] = RaggedBatch(buf=buf, B=B, T=T, L=L, hop=hop_length, n_mels=n_mels, lens=ship_all, mel_len_host=true_all,
                                    wav_off=wav_off, mel_off=mel_off, logd_off=logd_off, seg=seg,
                                    gather_bytes=4 * sum(x_[2] - (x_[1] if shared_out else 0) for x_ in seg[1:]),
                                    scatter_bytes=scatter_bytes,
                                    events=events if events is not None else {})
            st["mine"] = buf[:st["my_size"]]
            # host offsets of every rank's waveform segment: rank-major — contiguous, or fixed slots in a shared window
            ho, acc = [], 0
            for r_, (_, w, _) in enumerate(seg):
                ho.append(r_ * slot if shared_out else acc)
                acc += w
            st["host_seg"] = ho
            if host_t is not None:
                if not shared_out and host_t.numel() < acc:
                    raise ValueError(f"host_out holds {host_t.numel()} samples, the gathered waveforms need {acc}")
                st["out"].host = host_t
                st["out"].host_wav_off = [ho[r_] + st["out"].wav_off[i] - seg[r_][0]
                                          for r_, blk in enumerate(partition(B, world)) for i in blk]
                st["out"].wav_on_device = not (shared_out and multi)
        else:
            st["mine"] = torch.empty(st["my_size"], dtype=torch.float32, device=dev)
        if n_mine > 0:
            lens_d = mel_len.to(torch.int64) if not tails_padded else torch.full((n_mine,), L, dtype=torch.int64, device=dev)
            start = torch.cumsum(lens_d, 0) - lens_d
            st["lens_d"], st["w_offs"], st["m_offs"] = lens_d, start * hop_length, start * n_mels + F * hop_length
            acc, wl, ml = 0, [], []
            for v in st["my_ship"]:
                wl.append(acc * hop_length)
                ml.append(F * hop_length + acc * n_mels)
                acc += v
            st["wl"], st["ml"], st["cum"] = wl, ml, wl + [acc * hop_length]

    def ship(parts):
        """One grouped NCCL exchange: rank 0 receives ``parts`` = [(rank, start, end)] element ranges of every peer's buffer
        in place; a peer sends its own range.  Then (rank 0, host_out) the device-to-host copies of the waveform ranges.
        Runs on the side stream when there is one, so the next group's kernels overlap it."""
        def body():
            ops = []
            if multi:
                for (r, a0, a1, is_wav) in parts:
                    if a1 <= a0 or (shared_out and is_wav):
                        continue
                    if rank == 0 and r != 0:
                        s0 = st["out"].seg[r][0]
                        ops.append(dist.P2POp(dist.irecv, st["out"].buf[s0 + a0: s0 + a1],
                                              r if group is None else dist.get_global_rank(group, r), group))
                    elif rank == r and r != 0:
                        ops.append(dist.P2POp(dist.isend, st["mine"][a0:a1], root, group))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            if shared_out:                             # every rank: its own waveforms, over its own PCIe link
                for (r, a0, a1, is_wav) in parts:
                    if is_wav and r == rank and a1 > a0:
                        h0 = st["host_base"]
                        host_t[h0 + a0: h0 + a1].copy_(st["mine"][a0:a1], non_blocking=True)
            elif rank == 0 and host_t is not None:
                for (r, a0, a1, is_wav) in parts:
                    if is_wav and a1 > a0:
                        s0, h0 = st["out"].seg[r][0], st["host_seg"][r]
                        host_t[h0 + a0: h0 + a1].copy_(st["out"].buf[s0 + a0: s0 + a1], non_blocking=True)
        if side is None:
            body()
        else:
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(side):
                side.wait_event(ev)
                body()

    def peer_wav_range(r, i):
        """Element range (inside rank r's segment) of the waveforms of rank r's i-th vocoder group, from the lengths rank 0
        holds; None when rank r has no such group."""
        gb = group_bounds(counts[r], G) if counts[r] else []
        if i >= len(gb):
            return None
        base = sum(counts[:r])
        ship_all = st["out"].lens
        a0 = sum(ship_all[base: base + gb[i][0]]) * hop_length
        a1 = sum(ship_all[base: base + gb[i][1]]) * hop_length
        return a0, a1

    def on_group(i, g0, g1, wav, mel, mel_len):
        if "mine" not in st:
            prepare(mel.shape[2], mel_len)
        if g1 > g0:
            _ragged_copy(True, wav[g0:g1].to(torch.float32).contiguous(), st["mine"], st["lens_d"][g0:g1], st["w_offs"][g0:g1],
                         st["my_ship"][g0:g1], st["wl"][g0:g1], 1, hop_length)
        if shared_out:
            parts = [(rank, st["cum"][g0], st["cum"][g1], True)]
        elif rank == 0:
            parts = []
            for r in range(world):
                rng = peer_wav_range(r, i)
                if rng:
                    parts.append((r, rng[0], rng[1], True))
        else:
            parts = [(rank, st["cum"][g0], st["cum"][g1], True)]
        ship(parts)
        st["groups_done"] += 1

    kw = dict(vocoder_groups=G, on_group=on_group) if (G != 1) else {}
    if n_mine > 0:
        wav, mel, mel_len, logd = model(xs, force_duration=has_dur, pad_to=pad_to, zero_padded_mel=bool(zero_pad), **kw)
        L = mel.shape[2]
        mark("forward")
        if st["groups_done"] == 0:                     # the model vocoded in one piece (or ignores the group protocol)
            for i, (g0, g1) in enumerate(group_bounds(n_mine, G)):
                on_group(i, g0, g1, wav, mel, mel_len)
        # the small tail of the segment: valid mel frames + log-durations
        F = st["F"]
        _ragged_copy(True, mel.to(torch.float32).contiguous(), st["mine"], st["lens_d"], st["m_offs"], st["my_ship"], st["ml"],
                     n_mels, 1)
        d0 = F * (hop_length + n_mels)
        st["mine"][d0: d0 + n_mine * T].view(n_mine, T).copy_(logd)
    else:                                              # more ranks than utterances: still part of every collective
        if not has_dur and multi:
            _exchange_lengths({"mine": []}, nb, world, dev, group)
        mark("forward")
        prepare(0, None)
    mark("result_packed")
    if rank == 0:
        # groups the other ranks have but rank 0's own loop did not reach (it has at least as many, so: none) + the tails
        tails_parts = []
        for r in range(world):
            Fr = st["out"].seg[r][1]
            tails_parts.append((r, Fr, st["out"].seg[r][2], False))
        ship(tails_parts)
    elif n_mine > 0:
        ship([(rank, st["F"] * hop_length, st["my_size"], False)])
    if side is not None:
        torch.cuda.current_stream(dev).wait_stream(side)
    mark("gathered")
    if rank != 0:
        return None
    out = st["out"]
    return out if ragged else out.padded()


def _exchange_lengths(lens_box: dict, nb: int, world: int, dev, group) -> int:
    """Predicted durations: one all-gather of the per-utterance frame counts (control plane) -> the global frame count;
    rank 0 keeps all counts to size the gathered buffer."""
    mine = lens_box["mine"]
    t = torch.full((max(nb, 1),), -1, dtype=torch.int64)
    if mine:
        t[: len(mine)] = torch.tensor(mine, dtype=torch.int64)
    t = t.to(dev)
    allt = torch.empty((world * max(nb, 1),), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allt, t, group=group)
    vals = allt.cpu().view(world, max(nb, 1)).tolist()
    lens_box["all"] = [v for row in vals for v in row if v >= 0]
    return max([0] + lens_box["all"])


def mixed_language_forward(models: Mapping[str, Callable], x: Mapping[str, torch.Tensor] | None, lang: Sequence[str] | None,
                           force_duration: bool = False, sharded: bool = False, group=None,
                           device: torch.device | str | None = None, hop_length: int = 256, n_mels: int = 80):
    """BASELINE config 4's mixed EN/DE batch: utterance i is synthesised by ``models[lang[i]]``.

    The reference ties one checkpoint to one language (``ZeroVoxTTS(language=...)``, synthesize.py:48-100), so a mixed
    batch is one batch per weight set there too: utterances are grouped by the *model object* their tag maps to (two tags
    may share one weight set), every group runs as a batch of its own — through :func:`sharded_forward` when ``sharded``
    (rank 0 holds ``x`` and ``lang``; the other ranks pass ``None`` and learn the group plan from one small object
    broadcast) — and the results are merged back in the original order, zero-padded to the longest group.

    Returns ``(wav [B, L*hop], mel [B, n_mels, L], mel_len int64 [B], log_duration [B, T])`` (on rank 0 when sharded,
    ``None`` elsewhere).
    """
    rank = dist.get_rank(group) if sharded else 0
    plan = None
    if rank == 0:
        assert x is not None and lang is not None, "rank 0 must hold the batch and its language tags"
        B = x["phoneme"].shape[0]
        if len(lang) != B:
            raise ValueError(f"{len(lang)} language tags for {B} utterances")
        unknown = sorted({t for t in lang if t not in models})
        if unknown:
            raise KeyError(f"no model for language tag(s) {unknown}")
        first_tag, members = {}, {}
        for i, t in enumerate(lang):                      # group by weight set, in order of first appearance
            key = first_tag.setdefault(id(models[t]), t)
            members.setdefault(key, []).append(i)
        plan = [(k, v) for k, v in members.items()]
    if sharded:
        box = [plan]
        dist.broadcast_object_list(box, src=0, group=group)
        plan = box[0]

    results = []
    for tag, idx in plan:
        xg = None
        if rank == 0:
            sel = torch.as_tensor(idx, dtype=torch.long)
            xg = {k: v[sel.to(v.device)] for k, v in x.items() if isinstance(v, torch.Tensor)}
        if sharded:
            out = sharded_forward(models[tag], xg, force_duration=force_duration, group=group, device=device,
                                  hop_length=hop_length, n_mels=n_mels)
        else:
            out = models[tag](xg, force_duration=force_duration)
        results.append(out)
    if rank != 0:
        return None

    B, T = x["phoneme"].shape
    L = max(r[1].shape[2] for r in results)
    dev = results[0][0].device
    wav = torch.zeros((B, L * hop_length), dtype=torch.float32, device=dev)
    mel = torch.zeros((B, results[0][1].shape[1], L), dtype=torch.float32, device=dev)
    mel_len = torch.zeros((B,), dtype=torch.int64, device=dev)
    logd = torch.zeros((B, T), dtype=torch.float32, device=dev)
    for (tag, idx), (w, m, ml, ld) in zip(plan, results):
        sel = torch.as_tensor(idx, dtype=torch.long, device=dev)
        wav[sel, : w.shape[1]] = w.to(torch.float32)
        mel[sel, :, : m.shape[2]] = m.to(torch.float32)
        mel_len[sel] = ml.to(torch.int64)
        logd[sel] = ld.to(torch.float32)
    return wav, mel, mel_len, logd
