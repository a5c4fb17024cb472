"""Batch sharding of one global batch across the GPUs of a box (BASELINE config 4).

Utterances are independent in eval mode, so the data path needs no collective: rank 0 packs every rank's share of the
inputs into ONE byte buffer per rank and scatters it; every rank runs ``ZeroVox.forward`` on its shard; the results are
packed into one byte buffer per rank and gathered on rank 0.  ``torch.distributed`` is the plumbing (NCCL over
NVLink/NVSwitch on GPUs; the same code runs over gloo with CPU tensors, which is how the host logic is tested).

Exactness against an unsharded run.  The reference's batch-composition quirks make an utterance's tail
depend on the batch's padded lengths: every shard therefore keeps the *global* phoneme length T (inputs are never
trimmed) and decodes / vocodes at the *global* frame count ``L_pad`` (``pad_to``).  With forced durations rank 0 knows
``L_pad`` up front and ships it in the header; with predicted durations it is one 8-byte MAX all-reduce — the only
other collective, and control-plane only.
"""
from __future__ import annotations

from typing import Callable, Mapping, Sequence

import torch
import torch.distributed as dist

_HEADER = 8                                            # int64 words


def partition(lengths: Sequence[int], world: int) -> list[list[int]]:
    """Longest-first round-robin deal (phoneme count is a proxy for frames ~ 6*T): rank r gets utterances
    order[r::world]; every rank receives ceil(B/world) or floor(B/world) utterances with similar total length."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    return [order[r::world] for r in range(world)]


def _pack_inputs(x: Mapping[str, torch.Tensor], idx: Sequence[int], nb: int, has_mask: bool, has_dur: bool):
    """Byte image of one rank's shard, padded to nb utterances (padding rows repeat the shard's first utterance; they are
    dropped again by the valid count)."""
    sel = list(idx) + [idx[0] if idx else 0] * (nb - len(idx))
    sel_t = torch.as_tensor(sel, dtype=torch.long)
    parts = []
    for k in ("phoneme", "puncts"):
        parts.append(x[k].cpu().to(torch.int32)[sel_t].contiguous().view(torch.uint8).reshape(-1))
    if has_dur:
        parts.append(x["duration"].cpu().to(torch.int32)[sel_t].contiguous().view(torch.uint8).reshape(-1))
    parts.append(x["ref_mel"].cpu().to(torch.float32)[sel_t].contiguous().view(torch.uint8).reshape(-1))
    if has_mask:   # byte-sized rows last so that every wider view stays aligned
        parts.append(x["phoneme_mask"].cpu().to(torch.uint8)[sel_t].contiguous().reshape(-1))
    return torch.cat(parts)


def _unpack_inputs(buf: torch.Tensor, nb: int, T: int, T_ref: int, n_mels: int, has_mask: bool, has_dur: bool):
    out, off = {}, 0

    def take(nbytes):
        nonlocal off
        t = buf[off:off + nbytes]
        off += nbytes
        return t

    out["phoneme"] = take(nb * T * 4).view(torch.int32).reshape(nb, T)
    out["puncts"] = take(nb * T * 4).view(torch.int32).reshape(nb, T)
    if has_dur:
        out["duration"] = take(nb * T * 4).view(torch.int32).reshape(nb, T)
    out["ref_mel"] = take(nb * T_ref * n_mels * 4).view(torch.float32).reshape(nb, T_ref, n_mels)
    if has_mask:
        out["phoneme_mask"] = take(nb * T).reshape(nb, T).to(torch.bool)
    return out


def _input_bytes(nb, T, T_ref, n_mels, has_mask, has_dur):
    n = nb * T * 4 * (3 if has_dur else 2) + (nb * T if has_mask else 0) + nb * T_ref * n_mels * 4
    return (n + 15) // 16 * 16


def sharded_forward(model: Callable, x: Mapping[str, torch.Tensor] | None, force_duration: bool = False,
                    group=None, device: torch.device | str | None = None, hop_length: int = 256, n_mels: int = 80):
    """Run ``model(x_shard, force_duration=..., pad_to=L_pad)`` on every rank of ``group`` for the global batch ``x`` held
    by rank 0 (other ranks pass ``x=None``).  ``model`` is a ``ZeroVox`` (or any callable with that signature returning
    ``(wav [n, L*hop], mel [n, n_mels, L], mel_len int64 [n], log_duration [n, T])``).

    Returns on rank 0 the tuple of ``ZeroVox.forward`` for the whole batch in the original utterance order, padded to the
    global L_pad; ``None`` on the other ranks.  Collectives: one scatter (inputs), one gather (results), plus a single
    8-byte MAX all-reduce of the frame count when durations are predicted.
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device(device) if device is not None else (
        torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu"))

    # ---- header: shapes + flags + L_pad hint, broadcast as 8 int64 ---------------------------------------------
    hdr = torch.zeros(_HEADER, dtype=torch.int64)
    parts = None
    if rank == 0:
        assert x is not None, "rank 0 must hold the global batch"
        B, T = x["phoneme"].shape
        T_ref = x["ref_mel"].shape[1]
        has_mask, has_dur = "phoneme_mask" in x, bool(force_duration) and "duration" in x
        if has_mask:
            lengths = (~x["phoneme_mask"].cpu().to(torch.bool)).sum(1).tolist()
        else:
            lengths = [T] * B
        parts = partition(lengths, world)
        nb = max(len(p) for p in parts)
        L_hint = int(x["duration"].cpu().clamp(min=0).sum(1).max()) if has_dur else -1
        hdr[:] = torch.tensor([B, T, T_ref, nb, int(has_mask), int(has_dur), L_hint, x["ref_mel"].shape[2]])
    hdr = hdr.to(dev)
    dist.broadcast(hdr, src=0, group=group)
    B, T, T_ref, nb, has_mask, has_dur, L_hint, n_mels_in = (int(v) for v in hdr.cpu().tolist())
    has_mask, has_dur = bool(has_mask), bool(has_dur)

    # ---- ONE scatter of the packed inputs ---------------------------------------------------------------------
    nbytes = _input_bytes(nb, T, T_ref, n_mels_in, has_mask, has_dur)
    recv = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    send = None
    if rank == 0:
        send = []
        for r in range(world):
            b = _pack_inputs(x, parts[r], nb, has_mask, has_dur)
            pad = torch.zeros(nbytes, dtype=torch.uint8)
            pad[: b.numel()] = b
            send.append(pad.to(dev))
    dist.scatter(recv, send, src=0, group=group)
    xs = _unpack_inputs(recv, nb, T, T_ref, n_mels_in, has_mask, has_dur)

    # ---- forward on the shard, at the global frame count ------------------------------------------------------
    def l_pad(local_lmax: int) -> int:
        if L_hint >= 0:
            return max(L_hint, local_lmax)
        t = torch.tensor([local_lmax], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return int(t.item())

    wav, mel, mel_len, logd = model(xs, force_duration=has_dur, pad_to=l_pad)
    L = mel.shape[2]

    # ---- ONE gather of the packed results ---------------------------------------------------------------------
    f32 = torch.cat([wav.reshape(nb, -1), mel.reshape(nb, -1), logd.reshape(nb, -1)], dim=1).to(torch.float32).contiguous()
    out = torch.cat([mel_len.to(torch.int64).contiguous().view(torch.uint8).reshape(-1), f32.view(torch.uint8).reshape(-1)])
    gathered = [torch.empty_like(out) for _ in range(world)] if rank == 0 else None
    dist.gather(out, gathered, dst=0, group=group)
    if rank != 0:
        return None

    W = L * hop_length
    wav_g = torch.empty((B, W), dtype=torch.float32, device=dev)
    mel_g = torch.empty((B, n_mels, L), dtype=torch.float32, device=dev)
    len_g = torch.empty((B,), dtype=torch.int64, device=dev)
    logd_g = torch.empty((B, T), dtype=torch.float32, device=dev)
    row = W + n_mels * L + T
    for r in range(world):
        idx = torch.as_tensor(parts[r], dtype=torch.long, device=dev)
        n = len(parts[r])
        if n == 0:
            continue
        f = gathered[r][nb * 8:].view(torch.float32).reshape(nb, row)[:n]
        wav_g[idx] = f[:, :W]
        mel_g[idx] = f[:, W:W + n_mels * L].reshape(n, n_mels, L)
        logd_g[idx] = f[:, W + n_mels * L:]
        len_g[idx] = gathered[r][: nb * 8].view(torch.int64)[:n]
    return wav_g, mel_g, len_g, logd_g


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
