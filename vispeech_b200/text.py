"""Text -> phoneme-id bridge in front of `SynthesizerTrn.infer` (first half).

What the reference does: `text/__init__.py:9-17` (`cleaned_text_to_sequence`: one table look-up per phoneme symbol) on top of
`text/symbols.py:1-22` (the 519-entry symbol table: "_" + 401 Mandarin initials/finals with tone + 42 Japanese + 69 ARPAbet
+ 6 punctuation / pause marks), fed by `text/cleaner.py:34-64` (`text_to_phones`: language-tagged G2P front ends built on
pypinyin / jieba / g2pM / g2p_en / pyopenjtalk, none of which exist in this image - SURVEY.md 8c).

Here:
  * the symbol table, rebuilt from its structure (checked against the reference's list in tests and by a digest);
  * `cleaned_text_to_sequence` / `sequence_to_cleaned_text` with the reference's semantics (KeyError on unknown symbols);
  * `TextFrontend`: text -> ids with an LRU phoneme cache (serving sees the same sentences again and again; G2P is the
    slow host step in front of a 4 ms GPU call).  The G2P itself is pluggable: by default the reference's own
    `text.cleaner.text_to_phones` when its package is importable (`sys.path` holds the reference checkout and its
    dependencies are installed), otherwise only phoneme-level input is accepted: whitespace-separated symbols of the
    table, e.g. "n i3 h ao3 sp" - the format of the reference's filelists (`filelists/*.list`, third column).
"""
from __future__ import annotations

import hashlib
import threading
from collections import OrderedDict
from typing import Callable, Dict, Iterable, List, Optional, Sequence

# ---- symbol table (reference text/symbols.py:1-22) -------------------------------------------------------------------
_PAD = "_"
_PUNCT = ["!", "?", "…", ",", ".", "sp"]
_ZH_INITIALS = "b c ch d f g h j k l m n p q r s sh t x z zh".split()
_ZH_FINALS = ("a ai an ang ao e ei en eng i ia ian iang iao ie ii iii in ing io iong iou o ong ou u ua uai uan uang uei uen "
              "ueng uo v van ve vn").split()
# japanese: common phones first, then the ones only the JA front end emits (order is part of the id assignment)
_JA = ("ts. f. sh. ry. py. h. p. N. a. m. w. ky. n. d. j. cl. ny. z. o. y. t. u. r. pau ch. e. b. k. g. s. i. "
       "gy. my. hy. br by. v. ty. xx. U. I. dy.").split()
_EN_VOWELS = "AA AE AH AO AW AY EH ER EY IH IY OW OY UH UW".split()
_EN_CONSONANTS = "B CH D DH F G HH JH K L M N NG P R S SH T TH V W Y Z ZH".split()


def _build_symbols() -> List[str]:
    # every final exists plain and with erhua ("r"), each in tones 1-5; the reference lists them in sorted order
    zh = sorted(_ZH_INITIALS + [f + r + t for f in _ZH_FINALS for r in ("", "r") for t in "12345"])
    en = sorted(_EN_CONSONANTS + [v + s for v in _EN_VOWELS for s in "012"])
    return [_PAD] + zh + _JA + en + _PUNCT


symbols: List[str] = _build_symbols()
SYMBOLS_SHA256 = "b8a169c52a4dd14445a9fc644fa05619ae5c3f712587175e8349705e750b5955"   # of "\n".join(reference symbols)
if hashlib.sha256("\n".join(symbols).encode()).hexdigest() != SYMBOLS_SHA256 or len(symbols) != 519:
    raise ImportError("vispeech_b200.text: the rebuilt symbol table differs from the reference's (text/symbols.py)")

_symbol_to_id: Dict[str, int] = {s: i for i, s in enumerate(symbols)}
_id_to_symbol: Dict[int, str] = {i: s for i, s in enumerate(symbols)}


def cleaned_text_to_sequence(cleaned_text: Iterable[str]) -> List[int]:
    """reference text/__init__.py:9-17: a sequence of phoneme symbols -> ids; unknown symbols raise KeyError."""
    return [_symbol_to_id[symbol] for symbol in cleaned_text]


def sequence_to_cleaned_text(sequence: Iterable[int]) -> List[str]:
    return [_id_to_symbol[int(i)] for i in sequence]


# ---- G2P plug-in + cache ---------------------------------------------------------------------------------------------
_REPLACE = {"-": "sp", "--": "sp"}           # text/cleaner.py:11-13


def remove_invalid_phonemes(phones: Sequence[str]) -> List[str]:
    """text/cleaner.py:23-32: map "-" / "--" to "sp" and drop what the table does not know."""
    out = []
    for ph in phones:
        ph = _REPLACE.get(ph, ph)
        if ph in _symbol_to_id:
            out.append(ph)
    return out


def phones_from_phoneme_string(text: str) -> List[str]:
    """Phoneme-level input: whitespace-separated symbols.  Unknown tokens are an error here (a typo must not silently
    vanish the way unknown G2P output does in the reference)."""
    toks = [_REPLACE.get(t, t) for t in text.split()]
    bad = [t for t in toks if t not in _symbol_to_id]
    if bad:
        raise ValueError("not phoneme symbols of text/symbols.py: %s" % ", ".join(sorted(set(bad))[:8]))
    return toks


def reference_g2p() -> Optional[Callable[[str], List[str]]]:
    """The reference's own `text.cleaner.text_to_phones` if it can be imported in this process, else None."""
    try:
        from text.cleaner import text_to_phones          # noqa: the reference package, when on sys.path with its deps
        return text_to_phones
    except Exception:
        return None


class TextFrontend:
    """text -> phoneme ids with an LRU cache.

    `g2p`: callable text -> list of phoneme symbols.  None = the reference cleaner when importable; if it is not, only
    phoneme strings are accepted.  `append_period`: the serving route of the reference appends "。" to every request
    (inference_api.py:17) before G2P - kept for raw text, not applied to phoneme strings."""

    def __init__(self, g2p: Optional[Callable[[str], List[str]]] = None, cache_size: int = 4096, append_period: bool = True):
        self.g2p = g2p if g2p is not None else reference_g2p()
        self.cache_size, self.append_period = int(cache_size), append_period
        self._cache: "OrderedDict[str, tuple]" = OrderedDict()
        self._lock = threading.Lock()
        self.hits = self.misses = 0

    def _is_phoneme_string(self, text: str) -> bool:
        toks = text.split()
        return len(toks) > 0 and all(_REPLACE.get(t, t) in _symbol_to_id for t in toks)

    def text_to_phones(self, text: str) -> List[str]:
        if self._is_phoneme_string(text):
            return phones_from_phoneme_string(text)
        if self.g2p is None:
            raise ValueError("raw text needs a G2P front end (the reference's text.cleaner is not importable here); "
                             "pass whitespace-separated phoneme symbols instead")
        return remove_invalid_phonemes(self.g2p(text + ("。" if self.append_period else "")))

    def text_to_sequence(self, text: str) -> List[int]:
        with self._lock:
            hit = self._cache.get(text)
            if hit is not None:
                self._cache.move_to_end(text)
                self.hits += 1
                return list(hit)
        ids = tuple(cleaned_text_to_sequence(self.text_to_phones(text)))     # G2P runs outside the lock
        with self._lock:
            self.misses += 1
            self._cache[text] = ids
            self._cache.move_to_end(text)
            while len(self._cache) > self.cache_size:
                self._cache.popitem(last=False)
        return list(ids)

    __call__ = text_to_sequence
