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

from zerovox_b200.parallel import mixed_language_forward, partition, sharded_forward  # noqa: E402

HOP, NMEL = 4, 6


def fake_model(x, force_duration=False, pad_to=None):
    """Per-utterance deterministic function of the inputs (so sharding must not change it)."""
    ph, pu = x["phoneme"].long(), x["puncts"].long()
    n, T = ph.shape
    valid = ~x["phoneme_mask"] if "phoneme_mask" in x else torch.ones(n, T, dtype=torch.bool)
    dur = x["duration"].long().clamp(min=0) if force_duration else ((ph + pu) % 5 + 1) * valid
    mel_len = dur.sum(1)
    L = int(mel_len.max())
    if pad_to is not None:
        L = max(L, int(pad_to(L) if callable(pad_to) else pad_to))
    base = x["ref_mel"].sum(dim=(1, 2))[:, None] + (ph * valid).float().sum(1, keepdim=True)
    t = torch.arange(L)[None, :].float()
    mel = (base[:, :, None] + torch.arange(NMEL)[None, :, None].float() * 0.5 + t[:, None, :]) * (t < mel_len[:, None])[:, None, :]
    tw = torch.arange(L * HOP)[None, :].float()
    wav = torch.sin(base + tw * 0.01) * (tw < (mel_len * HOP)[:, None])
    logd = torch.log1p(dur.float())
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


def fake_model_de(x, force_duration=False, pad_to=None):
    """A second 'weight set': same signature, different function of the inputs (and longer utterances)."""
    wav, mel, mel_len, logd = fake_model(dict(x, puncts=x["puncts"] + 1), force_duration=force_duration, pad_to=pad_to)
    return -wav, mel + 100.0 * (mel != 0), mel_len, logd + 1.0


def check_mixed(out, x, lang, models, forced):
    """Every utterance must equal what its own model gives for the batch of its own weight set."""
    wav, mel, mel_len, logd = out
    B = x["phoneme"].shape[0]
    assert wav.shape[0] == mel.shape[0] == B and wav.shape[1] == mel.shape[2] * HOP
    groups = {}
    for i, t in enumerate(lang):
        groups.setdefault(id(models[t]), (models[t], []))[1].append(i)
    for model, idx in groups.values():
        sel = torch.tensor(idx)
        rw, rm, rl, rd = model({k: v[sel] for k, v in x.items()}, force_duration=forced)
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
                check_mixed(out, x, lang, models, forced)
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
        for (B, T, T_ref, seed, ragged, forced) in cases:
            x = make_batch(B, T, T_ref, seed, ragged, forced) if rank == 0 else None
            out = sharded_forward(fake_model, x, force_duration=forced, device="cpu", hop_length=HOP, n_mels=NMEL)
            if rank == 0:
                ref = fake_model(x, force_duration=forced)
                for name, a, b in zip(("wav", "mel", "mel_len", "log_duration"), out, ref):
                    assert a.shape == b.shape, (name, a.shape, b.shape)
                    assert torch.equal(a, b), f"{name} differs for case {(B, T, T_ref, seed, ragged, forced)}"
            else:
                assert out is None
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_partition_balances_and_covers():
    parts = partition([5, 9, 1, 7, 3, 8, 2], 3)
    assert sorted(i for p in parts for i in p) == list(range(7))
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert parts[0][0] == 1  # longest first
    assert partition([], 2) == [[], []]


@pytest.mark.timeout(180)
def test_sharded_forward_world2_gloo():
    cases = [(8, 12, 5, 0, False, True), (7, 9, 4, 1, True, True), (5, 6, 3, 2, True, False), (1, 4, 2, 3, False, False)]
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
