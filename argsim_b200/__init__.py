"""argsim_b200 -- B200-native drop-in for the one hot path of argsim/argsim: the sequence-VAE
ELBO training step (reference src/model.py + src/train.py) and the encoder-only mu embedding
(src/eval_embed*.py).  Host side mirrors the reference's Python call surface; all arithmetic
runs in hand-written sm_100a CUDA kernels behind the C ABI of include/argsim_b200.h.
There is no CPU fallback: importing works anywhere, computing needs a B200.
"""
from .util import Record, comp, partial, select  # noqa: F401
from .model import vAe, encode, decode, Session, Saver, pipe, global_variables_initializer, set_random_seed  # noqa: F401
