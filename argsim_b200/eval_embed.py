"""The two embedding flows of the reference's eval_embed*.py scripts over the hot path:

  * deterministic representation (src/eval_embed_reason.py:30-42, src/eval_embed.py): encode_capped every text,
    vpack to one eos-padded matrix, `model.z.eval({model.src: data[i:j]})` over partitions of 128 rows;
  * averaged representation with sentencepiece sampling (src/eval_embed_reason.py:47-54): 128 sampled
    segmentations of ONE text form one batch, z = mu of every row, averaged.

Both go through `sess.run(model.z, {model.src: ...})` -> argsim_embed; only the host glue lives here."""
import numpy as np

from . import util_sp
from .util_np import partition, vpack


def embed_texts(sess, model, vocab, texts, batch=128, cap=512):
    """(len(texts), dim_rep) float32: mu of every text's deterministic encoding (src/eval_embed_reason.py:33-39)."""
    data = [util_sp.encode_capped(vocab, t, cap=cap) for t in texts]
    data = vpack(data, (len(data), max(map(len, data))), vocab.eos_id(), np.int32)
    out = [sess.run(model.z, {model.src: data[i:j]}) for i, j in partition(len(data), batch)]
    return np.concatenate(out, axis=0)


def infer_avg(sess, model, vocab, sent, samples=128, cap=512):
    """(dim_rep,) float32: mean mu over `samples` sampled segmentations of one text (src/eval_embed_reason.py:47-51)."""
    bat = [util_sp.encode_capped_sample(vocab, sent, cap=cap) for _ in range(samples)]
    bat = vpack(bat, (len(bat), max(map(len, bat))), vocab.eos_id(), np.int32)
    z = sess.run(model.z, {model.src: bat})
    return np.mean(z, axis=0)


def embed_texts_sampled(sess, model, vocab, texts, samples=128, cap=512):
    """(len(texts), dim_rep): infer_avg of every text, stacked (src/eval_embed_reason.py:53-54)."""
    return np.stack([infer_avg(sess, model, vocab, t, samples, cap) for t in texts], axis=0)
