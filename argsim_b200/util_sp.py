"""sentencepiece helpers of the data feed, same names and results as reference src/util_sp.py
(tokenisation happens on the host BEFORE the hot path).

`nltk.sent_tokenize` (used only when a text exceeds `cap`, src/util_sp.py:55) is imported when
available; otherwise a regex splitter on sentence-final punctuation stands in -- that branch then
may cut at slightly different sentence boundaries than the reference (documented deviation)."""
import re

import numpy as np

from .util_np import vpack

try:  # pragma: no cover - nltk is absent from the build image
    from nltk.tokenize import sent_tokenize
except Exception:  # noqa: BLE001
    _SENT_END = re.compile(r'(?<=[.!?])["\')\]]*\s+')

    def sent_tokenize(text):
        return [s for s in _SENT_END.split(text) if s]


def load_spm(path):
    """loads a sentencepiece model file -> SentencePieceProcessor (src/util_sp.py:6-14)."""
    from sentencepiece import SentencePieceProcessor
    sp = SentencePieceProcessor()
    sp.load(path)
    return sp


def spm(name, path, size=8192, bos=2, eos=1, unk=0, coverage=0.9995):
    """trains a unigram sentencepiece model (src/util_sp.py:17-39): ids unk=0 eos=1 bos=2."""
    from sentencepiece import SentencePieceTrainer
    SentencePieceTrainer.train(
        '--model_prefix=%s --input=%s --vocab_size=%d --bos_id=%d --eos_id=%d --unk_id=%d '
        '--unk_surface=☹ --character_coverage=%s' % (name, path, size, bos, eos, unk, coverage))


def _fit(encode, text, cap):
    """shared shrink loop: whole text, else the longest sentence prefix whose encoding fits `cap`;
    returns (result, fitted?).  `encode` returns a list or a tuple of lists."""
    def length(r):
        return max(map(len, r)) if isinstance(r, tuple) else len(r)
    out = encode(text)
    if length(out) <= cap:
        return out, True
    sents = sent_tokenize(text)
    n = int(len(sents) * cap / length(out))
    while n > 0:
        out = encode(' '.join(sents[:n]))
        if length(out) <= cap:
            return out, True
        n -= 1
    return out, False


def encode_capped(vocab, text, cap=512):
    """ids of `text`, at most `cap` long: drops trailing sentences, truncates as a last resort
    (src/util_sp.py:42-63)."""
    ids, ok = _fit(vocab.encode_as_ids, text, cap)
    return ids if ok else ids[:cap]


def _sampler(vocab):
    return lambda x: vocab.sample_encode_as_ids(x, -1, 0.5)


def encode_capped_sample(vocab, text, cap=512):
    """like encode_capped with sampled segmentation (nbest=-1, alpha=0.5); falls back to the
    deterministic encoding when nothing fits (src/util_sp.py:66-87)."""
    ids, ok = _fit(_sampler(vocab), text, cap)
    return ids if ok else encode_capped(vocab, text, cap)


def encode_capped_sample_pair(vocab, text, cap=512):
    """two independent sampled segmentations of the same (possibly shortened) text
    (src/util_sp.py:90-111)."""
    enc = _sampler(vocab)
    pair, ok = _fit(lambda x: (enc(x), enc(x)), text, cap)
    if ok:
        return pair
    ids = encode_capped(vocab, text, cap)
    return ids, ids


def encode(vocab, sents, length=None, dtype=np.int32):
    """(len(sents), length) id matrix padded with eos (src/util_sp.py:114-124)."""
    rows = [vocab.encode_as_ids(s) for s in sents]
    if length is None:
        length = max(map(len, rows))
    return vpack(rows, (len(rows), length), vocab.eos_id(), dtype)


def decode(vocab, array):
    """text of one id row (cut at the first eos); a generator of texts for higher ranks
    (src/util_sp.py:127-140)."""
    array = np.asarray(array)
    if array.ndim > 1:
        return (decode(vocab, a) for a in array)
    ids = [int(i) for i in array]
    if vocab.eos_id() in ids:
        ids = ids[:ids.index(vocab.eos_id())]
    return vocab.decode_ids(ids)
