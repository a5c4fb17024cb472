"""Phone / punctuation id tables — same interface as zerovox/tts/symbols.py:1-49 (needed to size embeddings)."""
from __future__ import annotations


class Symbols:
    NO_PUNCT = "_NP_"

    def __init__(self, phones: str, puncts: str):
        self._phone2id = {p: i for i, p in enumerate(phones)}
        self._id2phone = {i: p for p, i in self._phone2id.items()}
        self._punct2id = {Symbols.NO_PUNCT: 0}
        self._punct2id.update({p: i for i, p in enumerate(puncts, start=1)})
        self._id2punct = {i: p for p, i in self._punct2id.items()}

    def is_phone(self, p) -> bool:
        return p in self._phone2id

    def encode_phone(self, phone) -> int:
        return self._phone2id[phone]

    def decode_phone(self, idx) -> str:
        return self._id2phone[idx]

    @property
    def num_phones(self) -> int:
        return len(self._phone2id)

    def is_punct(self, p) -> bool:
        return p in self._punct2id

    def encode_punct(self, punct) -> int:
        return self._punct2id[punct]

    def decode_punct(self, idx) -> str:
        return self._id2punct[idx]

    @property
    def num_puncts(self) -> int:
        return len(self._punct2id)
