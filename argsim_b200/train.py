#!/usr/bin/env python3
"""Train driver with the reference's flags, config schema and loop structure (src/train.py:4-121).

    python -m argsim_b200.train --rounds 1 --config config.json [--trial NAME] [--ckpt NAME] [--gpu 0]
                                [--seed 0] [--prefetch 16] [--sample] [--profile]

Differences from the reference, all outside the hot path: summaries go to `<log>/<trial>.jsonl`
(one JSON object per 250 steps with step_errt / step_loss_gen / step_loss_kld) instead of a
TensorBoard event file; checkpoints use the library's own container; `--profile` runs three warm-up
validation passes, then one between cudaProfilerStart/Stop with an NVTX range around the device program and an
NVTX mark per phase (argsim_profiler), and prints per-phase device times instead of writing a TF RunMetadata.  `--precision fp32` selects the validation mode.
Launched as `python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 -m argsim_b200.train ...`
the same loop trains data parallel on N GPUs (rows of every batch sharded, gradients all-reduced by the library).
"""
import argparse
import json
import os
import sys


def parse(argv=None):
    parser = argparse.ArgumentParser(description="""
trains a variational autoencoder on text.
logs validation statistics per 250 steps;
saves a checkpoint per 10000 steps aka one round;
the checkpoints are named after the trial name and the training round.
details are specified in the config file.
""", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('--trial', default="master", help="the trial name")
    parser.add_argument('--config', default="config.json", help="the config file")
    parser.add_argument('--ckpt', default=None, help="the checkpoint to resume")
    parser.add_argument('--gpu', default="0", help="the gpu to use")
    parser.add_argument('--seed', default=0, type=int, help="random seed")
    parser.add_argument('--rounds', default=0, type=int, help="numbers of training rounds")
    parser.add_argument('--prefetch', default=16, type=int, help="numbers of batches to prefetch")
    parser.add_argument('--sample', action='store_true', help="train with sentencepiece sampling")
    parser.add_argument('--profile', action='store_true', help="time one validation pass per phase")
    parser.add_argument('--precision', default='bf16', choices=('bf16', 'fp32'), help="bf16 tensor cores or fp32 validation mode")
    parser.add_argument('--steps-per-summary', default=250, type=int, help=argparse.SUPPRESS)
    parser.add_argument('--summaries-per-round', default=40, type=int, help=argparse.SUPPRESS)
    return parser.parse_args(argv)


def make_batch_fn(T, P, vocab, seed, kudo, encode_capped, encode_capped_sample_pair):
    """the reference's `batch` generator (src/train.py:54-68): exactly `size` rows per batch, src is tgt
    unless sampling; the size check precedes the append, so the stream never ends."""
    import numpy as np
    from .util_io import load_txt
    from .util_np import sample, vpack

    def batch(size=T.batch_train, path=P.train, seed=seed, max_len=T.max_len):
        eos = vocab.eos_id()
        raw = tuple(load_txt(path))
        enc = encode_capped_sample_pair if kudo else encode_capped

        def pac(arrs):
            return vpack(arrs, (size, max(map(len, arrs))), eos, np.int32)

        bat = []
        for i in sample(len(raw), seed):
            if size == len(bat):
                if kudo:
                    src, tgt = map(pac, zip(*bat))
                else:
                    src = tgt = pac(bat)
                yield src, tgt
                bat = []
            bat.append(enc(vocab, raw[i], cap=max_len))
    return batch


def main(argv=None):
    A = parse(argv)
    if not A.rounds and not A.profile:
        sys.exit("nothing to do")
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world == 1:
        os.environ['CUDA_VISIBLE_DEVICES'] = A.gpu   # src/train.py:14,26; under torch.distributed.run every rank takes GPU LOCAL_RANK

    import numpy as np
    from . import model as M
    from .util import Record
    from .util_io import pform, load_json
    from .util_np import partition
    from . import util_sp

    config = load_json(A.config)
    P, C, T = Record(config['paths']), Record(config['model']), Record(config['train'])
    M.set_random_seed(A.seed)
    vocab = util_sp.load_spm(P.vocab)
    valid = np.load(P.valid)

    model_valid = M.vAe('valid', **C)
    batch = make_batch_fn(T, P, vocab, A.seed, A.sample, util_sp.encode_capped, util_sp.encode_capped_sample_pair)
    src, tgt = M.pipe(batch, (np.int32, np.int32), prefetch=A.prefetch)
    model_train = M.vAe('train', src=src, tgt=tgt, **C)

    sess = M.Session(precision=A.precision)
    if A.sample:
        # sentencepiece's sampling RNG is per process and unseeded (src/util_sp.py:66-111): under data parallelism the
        # ranks' batch streams differ, so rank 0's batch is the batch (Session.run broadcasts it before sharding)
        sess.sync_batches = 'broadcast'
    saver = M.Saver()
    if A.ckpt:
        saver.restore(sess, pform(P.ckpt, A.ckpt))
    else:
        M.global_variables_initializer(sess)

    if A.profile:
        feed = {model_valid.src: valid[:32], model_valid.tgt: valid[:32]}
        for _ in range(3):
            sess.run(model_valid.loss, feed)
        sess.handle.profiler(True)    # cudaProfilerStart + NVTX ranges: `nsys/ncu --capture-range=cudaProfilerApi`
        sess.run(model_valid.loss, feed)
        sess.handle.profiler(False)
        print(json.dumps(dict(profile=sess.handle.last_timings())))
        if not A.rounds:
            sys.exit("profiling done")

    # data parallel (python -m torch.distributed.run --nproc-per-node N -m argsim_b200.train ...): every rank runs this
    # same loop on the same batch stream and trains on its rows of each batch (Session.run shards them); T.batch_train is
    # the GLOBAL batch. Rank 0 alone validates, logs and saves -- the weights are identical on every rank.
    main_rank = sess.rank == 0
    if main_rank:
        os.makedirs(os.path.expanduser(P.log), exist_ok=True)
        os.makedirs(os.path.expanduser(P.ckpt), exist_ok=True)
    log = open(pform(P.log, A.trial, '.jsonl'), 'a') if main_rank else None

    def summ(step, model=model_valid):
        if not main_rank:
            return
        parts = [sess.run((model.errt_samp, model.loss_gen_samp, model.loss_kld_samp),
                          {model.src: valid[i:j], model.tgt: valid[i:j]})
                 for i, j in partition(len(valid), T.batch_valid, discard=False)]
        errt, gen, kld = (float(np.mean(np.concatenate([np.ravel(p[k]) for p in parts]))) for k in range(3))
        log.write(json.dumps(dict(step=int(step), step_errt=errt, step_loss_gen=gen, step_loss_kld=kld)) + '\n')
        log.flush()

    try:
        from tqdm import tqdm
    except ImportError:
        def tqdm(x, **kw):
            return x
    step = sess.run(model_train.step)
    for _ in range(A.rounds):
        for _ in range(A.summaries_per_round):
            for _ in tqdm(range(A.steps_per_summary), ncols=70):
                sess.run(model_train.train_step)
            step = sess.run(model_train.step)
            summ(step)
        if main_rank:
            saver.save(sess, pform(P.ckpt, A.trial, step // 10000), write_meta_graph=False)
    if log:
        log.close()


if __name__ == '__main__':
    main()
