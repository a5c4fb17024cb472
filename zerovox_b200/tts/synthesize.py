"""Mirror of the inference-time part of zerovox/tts/synthesize.py: the ``ZeroVoxTTS`` methods that sit on either side of
the hot path — ``speaker_embed`` (synthesize.py:123-143), ``transcript2phonemids`` (145-190), ``text2phonemeids``
(192-213), ``tts_ex`` / ``tts`` (215-243) — with the GPU front-end and the native tokeniser underneath.

Not mirrored (outside the path, SURVEY.md §8f / DESIGN.md §9): audio file decoding (librosa.load; its `sr=` resampling IS
built: `speaker_embed(wav, sampling_rate=)`), the NeMo /
uroman text normaliser (pass any callable with ``normalize(text) -> (transcript, _)`` as ``normalizer``), the HuggingFace
download in ``load_model`` (local directory / cache layouts are read, nothing is fetched) and the torchinfo ``summary``.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from ..frontend import MelFrontend, Tokeniser
from .model import ZeroVox
from .symbols import Symbols


class ZeroVoxTTS:

    def __init__(self, language: str, syms: Symbols, checkpoint, meldec_model: str, hop_length: int, sampling_rate: int,
                 n_mel_channels: int, fft_size: int, win_length: int, mel_fmin: int, mel_fmax: int,
                 infer_device: str = "cuda", num_threads: int = -1, verbose: bool = False, *, model: ZeroVox | None = None,
                 normalizer=None):
        self._hop_length = hop_length
        self._infer_device = infer_device
        self._sampling_rate = sampling_rate
        self._language = language
        self._meldec_model = meldec_model
        self._fft_size = fft_size
        self._win_length = win_length
        self._num_mels = n_mel_channels
        self._mel_fmin = mel_fmin
        self._mel_fmax = mel_fmax
        self._verbose = verbose
        if model is None:      # synthesize.py:78-88
            model = ZeroVox.load_from_checkpoint(lang=language, meldec_model=meldec_model, sampling_rate=sampling_rate,
                                                 hop_length=hop_length, checkpoint_path=str(checkpoint),
                                                 infer_device=infer_device, map_location=torch.device("cpu"),
                                                 strict=False, verbose=verbose, betas=[0.9, 0.99], eps=1e-9)
        self._model = model.to(infer_device)
        self._model.eval()
        self._symbols = syms
        self._normalizer = normalizer
        phones = "".join(syms.decode_phone(i) for i in range(syms.num_phones))
        puncts = "".join(syms.decode_punct(i) for i in range(1, syms.num_puncts))
        self._tokeniser = Tokeniser(phones, puncts)
        self._frontend = MelFrontend(sampling_rate, fft_size, hop_length, win_length, n_mel_channels, mel_fmin, mel_fmax,
                                     device=infer_device)

    def speaker_embed(self, wav, sampling_rate=None):
        """synthesize.py:123-143: trim(top_db=40) -> log-mel -> `_spkemb`; the prompt goes to the GPU once and only the
        trimmed length (16 bytes) comes back before the style vector.  ``sampling_rate`` (extension): the rate of ``wav`` when it
        is not the model's — the conversion `librosa.load(..., sr=)` does in get_speakerref (synthesize.py:113-121; the packaged
        prompts are 24 kHz) then runs on the GPU first."""
        if not isinstance(wav, torch.Tensor):
            wav = torch.from_numpy(np.ascontiguousarray(wav, dtype=np.float32))
        wav = wav.to(self._infer_device)
        with torch.no_grad():
            if sampling_rate is not None and int(sampling_rate) != int(self._sampling_rate):
                wav = self._frontend.resample(wav, int(sampling_rate), int(self._sampling_rate))[0]
            x = self._frontend.speaker_prompt_mel(wav, top_db=40)
            return self._model._spkemb(x)

    def transcript2phonemids(self, transcript: str) -> tuple[list[int], list[int]]:
        return self._tokeniser.transcript2phonemids(transcript)

    def text2phonemeids(self, text: str) -> tuple[list[int], list[int]]:
        if self._normalizer is None:
            raise RuntimeError("text normalisation (NeMo + uroman, zerovox/tts/normalize.py) is outside this engine: "
                               "pass normalizer= to ZeroVoxTTS or call transcript2phonemids on normalised text")
        transcript_uroman, _ = self._normalizer.normalize(text)
        phone_ids, punct_ids = self.transcript2phonemids(transcript_uroman)
        if self._verbose:
            print(f"Raw Text Sequence: {text}")
            print(f"Normalized       : {transcript_uroman}")
            print(f"Phoneme IDs      : {phone_ids}")
            print(f"Punct IDs        : {punct_ids}")
        return phone_ids, punct_ids

    def tts_ex(self, text: str, spkemb, duration=None):
        text = text.strip()
        tstart_g2p = time.time()
        phone_ids, punct_ids = self.text2phonemeids(text)
        if not phone_ids:   # synthesize.py:221-222
            return (np.array([[0.0]], dtype=np.float32), np.array([[0]], dtype=np.int32), 0,
                    np.array([[0.0]], dtype=np.float32))
        tend_g2p = time.time()
        tstart_synth = time.time()
        with torch.no_grad():
            phoneme = torch.tensor([phone_ids], dtype=torch.int32).to(self._infer_device)
            puncts = torch.tensor([punct_ids], dtype=torch.int32).to(self._infer_device)
            duration = torch.tensor([duration], dtype=torch.int32).to(self._infer_device) if duration is not None else None
            wav, length, _, mel = self._model.inference_ex({"phoneme": phoneme, "puncts": puncts, "duration": duration},
                                                           style_embed=spkemb, force_duration=duration is not None)
            wav = wav.cpu().numpy()
        tend_synth = time.time()
        if self._verbose:
            print(f"tts timing stats: g2p={tend_g2p-tstart_g2p}s, synth={tend_synth-tstart_synth}s")
        return wav, phoneme, length, mel.cpu().detach().numpy()

    def tts(self, text: str, spkemb):
        wav, phoneme, length, _ = self.tts_ex(text=text, spkemb=spkemb)
        return wav, phoneme, length

    @classmethod
    def load_model(cls, modelpath, meldec_model, infer_device: str = "cuda", num_threads: int = -1, verbose: bool = False,
                   *, normalizer=None):
        """synthesize.py:275-328: ``modelpath`` is a model directory (``modelcfg.yaml`` + ``checkpoints/*.ckpt``, newest
        wins) or a model name resolved in the local cache layout ``$CACHED_PATH_ZEROVOX/model_repo/<name>/{modelcfg.yaml,
        checkpoint.pkl}`` (model.py:66-82).  Nothing is downloaded: a missing file raises FileNotFoundError.
        Returns ``(modelcfg, ZeroVoxTTS)`` like the reference."""
        import glob
        from pathlib import Path

        import yaml

        from .model import model_cache_path
        if os.path.isdir(modelpath):
            config_path = Path(modelpath) / "modelcfg.yaml"
            list_of_files = glob.glob(os.path.join(modelpath, "checkpoints/*.ckpt"))
            if not list_of_files:
                raise FileNotFoundError(f"no checkpoints/*.ckpt under {modelpath}")
            checkpoint = max(list_of_files, key=os.path.getctime)
        else:
            config_path = model_cache_path(str(modelpath), "modelcfg.yaml")
            checkpoint = model_cache_path(str(modelpath), "checkpoint.pkl")
        if verbose:
            print("synthesize: using config    : ", config_path)
            print("synthesize: using checkpoint: ", checkpoint)
        with open(config_path) as modelcfgf:
            modelcfg = yaml.load(modelcfgf, Loader=yaml.FullLoader)
        synth = cls(language=modelcfg["lang"][0],
                    syms=Symbols(phones=modelcfg["model"]["phones"], puncts=modelcfg["model"]["puncts"]),
                    checkpoint=checkpoint, meldec_model=str(meldec_model), hop_length=modelcfg["audio"]["hop_size"],
                    win_length=modelcfg["audio"]["win_length"], mel_fmin=modelcfg["audio"]["fmin"],
                    mel_fmax=modelcfg["audio"]["fmax"], sampling_rate=modelcfg["audio"]["sampling_rate"],
                    n_mel_channels=modelcfg["audio"]["num_mels"], fft_size=modelcfg["audio"]["fft_size"],
                    infer_device=infer_device, num_threads=num_threads, verbose=verbose, normalizer=normalizer)
        return modelcfg, synth

    @property
    def normalizer(self):
        return self._normalizer

    @property
    def meldec_model(self):
        return self._meldec_model
