"""Host logic of the batch-sharded forward (zerovox_b200/parallel.py) over gloo, world_size 2, CPU tensors.

The model is replaced by a deterministic stand-in with ZeroVox.forward's signature (the CUDA engine has no CPU path);
what is checked is the plumbing: partitioning, the single packed scatter, padding to the global frame count, the single
packed gather and the restoration of the original utterance order — for forced and for predicted durations, ragged
batches and batches smaller than the world size.
"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from zerovox_b200.parallel import (SharedHostBatch, SharedHostBuffer, mixed_language_forward, partition,  # noqa: E402
                                   sharded_forward)

HOP, NMEL = 4, 6


def fake_model(x, force_duration=False, pad_to=None, zero_padded_mel=None, vocoder_groups=None, on_group=None):
    """Per-utterance deterministic function of the inputs (so sharding must not change it).  Like the reference
    (model.py:283-285) the padded frames are zero-filled only when a mel mask exists and the batch has more than one
    utterance — otherwise they hold a length-dependent non-zero pattern, so a shard that decides this from its LOCAL batch
    size, or pads to its local frame count, is caught."""
    ph, pu = x["phoneme"].long(), x["puncts"].long()
    n, T = ph.shape
    valid = ~x["phoneme_mask"] if "phoneme_mask" in x else torch.ones(n, T, dtype=torch.bool)
    dur = x["duration"].long().clamp(min=0) if force_duration else ((ph + pu) % 5 + 1) * valid
    mel_len = dur.sum(1)
    L = int(mel_len.max())
    if pad_to is not None:
        L = max(L, int(pad_to(L, mel_len.tolist()) if callable(pad_to) else pad_to))
    zero = zero_padded_mel if zero_padded_mel is not None else (((not force_duration) or "mel_mask" in x) and n > 1)
    base = x["ref_mel"].sum(dim=(1, 2))[:, None] + (ph * valid).float().sum(1, keepdim=True)
    t = torch.arange(L)[None, :].float()
    inside = (t < mel_len[:, None])[:, None, :]
    mel = base[:, :, None] + torch.arange(NMEL)[None, :, None].float() * 0.5 + t[:, None, :]
    mel = torch.where(inside, mel, torch.zeros(()) if zero else (t[:, None, :] - L) * torch.ones(n, NMEL, 1))
    tw = torch.arange(L * HOP)[None, :].float()
    inside_w = tw < (mel_len * HOP)[:, None]
    wav = torch.where(inside_w, torch.sin(base + tw * 0.01), torch.zeros(()) if zero else torch.cos(tw - L * HOP) * torch.ones(n, 1))
    logd = torch.log1p(dur.float())
    if on_group is not None:
        from zerovox_b200.tts.model import group_bounds
        for i, (g0, g1) in enumerate(group_bounds(n, vocoder_groups)):
            on_group(i, g0, g1, wav, mel, mel_len)
    return wav, mel, mel_len, logd


def make_batch(B, T, T_ref, seed, ragged, forced):
    g = torch.Generator().manual_seed(seed)
    x = {"phoneme": torch.randint(1, 28, (B, T), generator=g, dtype=torch.int32),
         "puncts": torch.randint(0, 10, (B, T), generator=g, dtype=torch.int32),
         "ref_mel": torch.randn(B, T_ref, NMEL, generator=g)}
    if ragged:
        lens = torch.randint(1, T + 1, (B,), generator=g)
        lens[0] = T
        x["phoneme_mask"] = torch.arange(T)[None, :] >= lens[:, None]
    if forced:
        d = torch.randint(0, 7, (B, T), generator=g, dtype=torch.int32)
        if ragged:
            d = d * (~x["phoneme_mask"])
        x["duration"] = d
    return x


def fake_model_de(x, force_duration=False, pad_to=None, zero_padded_mel=None, **kw):
    """A second 'weight set': same signature, different function of the inputs (and longer utterances)."""
    wav, mel, mel_len, logd = fake_model(dict(x, puncts=x["puncts"] + 1), force_duration=force_duration, pad_to=pad_to,
                                         zero_padded_mel=True)
    return -wav, mel + 100.0 * (mel != 0), mel_len, logd + 1.0


def assert_valid_equal(out, ref, exact_tails):
    """out == ref on every utterance's own samples / frames; past them: equal when the tails were shipped, else zero."""
    wav, mel, mel_len, logd = out
    rw, rm, rl, rd = ref
    assert torch.equal(mel_len, rl) and torch.equal(logd, rd)
    assert wav.shape[1] == mel.shape[2] * HOP and mel.shape[2] >= rm.shape[2]
    if exact_tails:
        assert wav.shape == rw.shape and mel.shape == rm.shape
        assert torch.equal(wav, rw) and torch.equal(mel, rm)
        return
    for i, n in enumerate(rl.tolist()):
        assert torch.equal(mel[i, :, :n], rm[i, :, :n]) and not mel[i, :, n:].any()
        assert torch.equal(wav[i, : n * HOP], rw[i, : n * HOP]) and not wav[i, n * HOP:].any()


def check_mixed(out, x, lang, models, forced, sharded=False):
    """Every utterance must equal what its own model gives for the batch of its own weight set (unsharded: tails included,
    zeros past the group's own frame count; sharded: the valid parts, zeros past them)."""
    wav, mel, mel_len, logd = out
    B = x["phoneme"].shape[0]
    assert wav.shape[0] == mel.shape[0] == B and wav.shape[1] == mel.shape[2] * HOP
    groups = {}
    for i, t in enumerate(lang):
        groups.setdefault(id(models[t]), (models[t], []))[1].append(i)
    for model, idx in groups.values():
        sel = torch.tensor(idx)
        ref = model({k: v[sel] for k, v in x.items()}, force_duration=forced)
        if sharded:
            assert_valid_equal((wav[sel], mel[sel], mel_len[sel], logd[sel]), ref, exact_tails=False)
            continue
        rw, rm, rl, rd = ref
        L = rm.shape[2]
        assert torch.equal(mel_len[sel], rl) and torch.equal(logd[sel], rd)
        assert torch.equal(mel[sel][:, :, :L], rm) and not mel[sel][:, :, L:].any()
        assert torch.equal(wav[sel][:, : L * HOP], rw) and not wav[sel][:, L * HOP:].any()


def test_mixed_language_single_process():
    models = {"en": fake_model, "de": fake_model_de, "en-gb": fake_model}        # two tags share one weight set
    for B, forced, ragged in ((6, True, False), (7, False, True), (1, True, False)):
        x = make_batch(B, 9, 4, 20 + B, ragged, forced)
        lang = [("en", "de", "en-gb")[i % 3] for i in range(B)]
        out = mixed_language_forward(models, x, lang, force_duration=forced, hop_length=HOP, n_mels=NMEL)
        check_mixed(out, x, lang, models, forced)
    with pytest.raises(KeyError):
        mixed_language_forward(models, make_batch(2, 4, 2, 0, False, True), ["en", "fr"], force_duration=True)
    with pytest.raises(ValueError):
        mixed_language_forward(models, make_batch(2, 4, 2, 0, False, True), ["en"], force_duration=True)


def _worker_mixed(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        models = {"en": fake_model, "de": fake_model_de}
        for B, forced, ragged in ((9, True, True), (6, False, False), (3, True, False)):
            x = make_batch(B, 10, 4, 40 + B, ragged, forced) if rank == 0 else None
            lang = ["en" if i % 2 == 0 else "de" for i in range(B)] if rank == 0 else None      # alternating tags
            out = mixed_language_forward(models, x, lang, force_duration=forced, sharded=True, device="cpu",
                                         hop_length=HOP, n_mels=NMEL)
            if rank == 0:
                check_mixed(out, x, lang, models, forced, sharded=True)
            else:
                assert out is None
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_mixed_language_sharded_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_mixed, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(30)
    assert all(r[1] == "ok" for r in res), res


def _worker(rank, world, port, cases, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for (B, T, T_ref, seed, ragged, forced, mel_mask) in cases:
            x = make_batch(B, T, T_ref, seed, ragged, forced) if rank == 0 else None
            if rank == 0 and mel_mask:
                ml = x["duration"].long().clamp(min=0).sum(1)
                x["mel_mask"] = torch.arange(int(ml.max()))[None, :] >= ml[:, None]
            ref = fake_model(x, force_duration=forced) if rank == 0 else None
            for tails in ("valid", "padded"):
                out = sharded_forward(fake_model, x, force_duration=forced, device="cpu", hop_length=HOP, n_mels=NMEL,
                                      tails=tails)
                if rank == 0:
                    assert_valid_equal(out, ref, exact_tails=(tails == "padded"))
                else:
                    assert out is None
            # the ragged container itself, with the header known on every rank (no broadcast)
            Lh = int(x["duration"].clamp(min=0).sum(1).max()) if (rank == 0 and forced) else -1
            box = [[B, T, T_ref, NMEL, int(ragged), int(forced), int(mel_mask), Lh,
                    x["mel_mask"].shape[1] if (rank == 0 and mel_mask) else 0,
                    int(((not forced) or mel_mask) and B > 1), 0, 0]]
            dist.broadcast_object_list(box, src=0)
            host = torch.zeros(B * T * 7 * HOP + 64) if rank == 0 else None
            rb = sharded_forward(fake_model, x, force_duration=forced, device="cpu", hop_length=HOP, n_mels=NMEL,
                                 ragged=True, spec=box[0], vocoder_groups=(3 if B % 2 else -3), host_out=host)
            if rank == 0:
                o = 0
                for i, n in enumerate(ref[2].tolist()):     # host_out: the valid waveforms back to back, utterance order
                    assert torch.equal(host[o:o + n * HOP], ref[0][i, : n * HOP]), f"host_out utterance {i}"
                    o += n * HOP
                assert rb.B == B and len(rb.wav_segments()) <= world and rb.gather_bytes % 4 == 0
                for i, n in enumerate(ref[2].tolist()):
                    assert torch.equal(rb.wav(i), ref[0][i, : n * HOP]) and torch.equal(rb.mel(i), ref[1][i, :, :n])
                assert torch.equal(rb.log_duration(), ref[3])
            else:
                assert rb is None
            # host-to-host through shared host windows: every rank uploads its own block and writes its own waveforms
            shared_x = SharedHostBatch(x, register=False)
            win = SharedHostBuffer(4 * (world * (-(-B // world)) * T * 7 * HOP + 64), register=False)
            win.tensor().zero_()
            dist.barrier()
            rb = sharded_forward(fake_model, shared_x, force_duration=forced, device="cpu", hop_length=HOP, n_mels=NMEL,
                                 ragged=True, vocoder_groups=(2 if B % 2 else 1), host_out=win)
            if rank == 0:
                assert not rb.wav_on_device and rb.host is not None
                for i, n in enumerate(ref[2].tolist()):
                    assert torch.equal(rb.host_wav(i), ref[0][i, : n * HOP]), f"shared host window, utterance {i}"
                    assert torch.equal(rb.mel(i), ref[1][i, :, :n])
                assert torch.equal(rb.log_duration(), ref[3])
                with pytest.raises(RuntimeError):
                    rb.wav(0)
            else:
                assert rb is None
            dist.barrier()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_partition_consecutive_blocks():
    parts = partition(7, 3)
    assert [list(p) for p in parts] == [[0, 1, 2], [3, 4, 5], [6]]
    assert [list(p) for p in partition(2, 4)] == [[0], [1], [], []]
    assert [list(p) for p in partition(0, 2)] == [[], []]


def test_sharded_forward_single_process_matches_model():
    """world size 1, no process group: the same container API, nothing to scatter."""
    for forced in (True, False):
        x = make_batch(5, 7, 3, 11, True, forced)
        ref = fake_model(x, force_duration=forced)
        out = sharded_forward(fake_model, x, force_duration=forced, device="cpu", hop_length=HOP, n_mels=NMEL, tails="padded")
        assert_valid_equal(out, ref, exact_tails=True)


@pytest.mark.timeout(180)
def test_sharded_forward_world2_gloo():
    # (B, T, T_ref, seed, ragged, forced, mel_mask); B = 2 on 2 ranks: one utterance per shard, the zero-fill decision of
    # model.py:283-285 must come from the global batch; B = 1: a rank without work
    cases = [(8, 12, 5, 0, False, True, False), (7, 9, 4, 1, True, True, False), (5, 6, 3, 2, True, False, False),
             (1, 4, 2, 3, False, False, False), (2, 6, 3, 4, True, False, False), (2, 5, 3, 5, False, True, False),
             (2, 5, 3, 6, False, True, True), (6, 8, 3, 7, True, True, True)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cases, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(30)
    assert all(r[1] == "ok" for r in res), res
