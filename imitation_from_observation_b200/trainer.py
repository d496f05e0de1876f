"""``ModelTrainer`` — the supervised trainer of scripts/train_script.py:25-203, on the B200 path.

Same constructor arguments and the same inner loop (random (src, tgt) video pairing with time folded into the
batch, Adam lr 1e-4, loss / sim / r1 / r2 / nn_err every 4th step, validation every 40), written against the
``compat`` adapter so the body reads like the reference.  Differences, all outside the hot path:

  * demos come from a ``[nlen, N, H, W, 3]`` array (the trainer's own ``vdata_*.npy`` format, train_script.py:97)
    or the synthetic generator — the mp4 / imageio / ffmpeg front-end (:59-96) is out of scope;
  * ``inception=True`` trains ``model(strides, kernels, filters)`` on the Mixed_7c feature maps of a frozen
    InceptionV3 (train_script.py:100-115); pass ``inception_net=`` (an ``inception.InceptionV3`` with weights
    loaded) — the reference restores ``model/inception_v3.ckpt``, which does not ship;
  * `device_bank=True` (default for pixel models) keeps vdata in HBM (`data.DeviceBank`): the loop's indexing
    expressions are unchanged but return device batches assembled by `ctx_gather_frames` — no per-step host gather,
    float64->float32 cast or H2D copy; the batches are bit-identical to the host path's;
  * data parallelism lives in parallel.py (bucketed gradient all-reduce) and is wired up by bench.py; this
    class is the single-process trainer the reference has.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import numpy as np

from . import _lib, compat as tf, synthetic
from .models import MODEL_REGISTRY


class ModelTrainer:
    def __init__(self, idims, nvideos, ntrain, batch_size, model, nitr, save_every, nlen, nskip, rescale=True,
                 inception=False, strides=None, kernels=None, filters=None, data=None, seed=0, precision="mixed",
                 log: Optional[Callable[[str], None]] = print, basedir="model/", inception_net=None,
                 device_bank=True):
        self.idims = tuple(idims)
        self.nvideos, self.ntrain, self.batch_size = nvideos, ntrain, batch_size
        self.nitr, self.save_every, self.nlen, self.nskip = nitr, save_every, nlen, nskip
        self.rescale, self.inception = rescale, inception
        self.strides, self.kernels, self.filters = strides, kernels, filters
        self.inception_net = inception_net
        self.device_bank = bool(device_bank) and not inception
        if inception and inception_net is None:
            raise _lib.CtxError("inception=True needs inception_net= (an inception.InceptionV3 with weights loaded)")
        if model not in MODEL_REGISTRY:
            raise _lib.CtxError("unknown model %r" % (model,))
        self.model = MODEL_REGISTRY[model]
        self.data, self.seed, self.precision, self.log, self.basedir = data, seed, precision, log or (lambda s: None), basedir
        self.history = []

    def _load_data(self) -> np.ndarray:
        """vdata [nlen, N, H, W, 3] float in [-1, 1] (train_script.py:83-97)."""
        if self.data is None:
            bank = synthetic.video_bank(self.nlen, self.nvideos, self.idims[0], self.idims[1], seed=1234 + self.seed)
            return bank if self.inception else synthetic.to_float(bank)   # inception: uint8 frames (rescale=False)
        if isinstance(self.data, str) and not self.data.endswith(".npy"):
            # a glob of demo videos (train_script.py:58: "model/videos/*.mp4"): decode, pick frames, transform, save the bank
            import glob
            from . import data as D
            bank = D.videos_to_bank(sorted(glob.glob(self.data)), self.nlen, self.nskip, self.idims, self.rescale,
                                    self.nvideos, self.inception, np.random.RandomState(self.seed), self.log)
            os.makedirs(self.basedir, exist_ok=True)
            np.save(os.path.join(self.basedir, "vdata_strike%d" % bank.shape[1]), bank)          # :90
            self.log(str(bank.shape))
            return bank
        if isinstance(self.data, str):
            return np.load(self.data)
        return np.asarray(self.data)

    def train(self):
        nlen = self.nlen
        vdata = self._load_data()
        batch_size = self.batch_size
        tf.reset_default_graph()
        if self.inception:                                                                        # :100-115
            tfinput = tf.placeholder(tf.uint8, (3, batch_size) + self.idims + (3,), name="image")
            image_trans = tf.multiply(tf.subtract(tf.image.convert_image_dtype(tfinput, dtype=tf.float32), 0.5), 2.0)
            featreshape = tf.inception_mixed_7c(image_trans, self.inception_net)
            test = self.model(strides=self.strides, kernels=self.kernels, filters=self.filters)
            with tf.variable_scope("contextmodel"):
                tf.build_model(test, featreshape, precision=self.precision, seed=self.seed)
            if np.asarray(vdata).dtype != np.uint8:
                raise _lib.CtxError("inception=True expects uint8 frames (the reference stores them with rescale=False)")
        else:
            tfinput = tf.placeholder(tf.float32, (3, batch_size) + self.idims + (3,), name="x")   # :118
            test = self.model()
            with tf.variable_scope("contextmodel"):                                               # :120-121
                tf.build_model(test, tfinput, precision=self.precision, seed=self.seed)
        sess = tf.Session()
        if self.device_bank:
            from .data import DeviceBank
            vdata = DeviceBank(vdata)
        learning_rate = tf.placeholder(tf.float32, shape=[])
        train_vars = tf.get_collection(tf.GraphKeys.TRAINABLE_VARIABLES, "contextmodel")
        optimizer = tf.train.AdamOptimizer(learning_rate).minimize(test.loss, var_list=train_vars)   # :128
        sess.run(tf.global_variables_initializer())
        allloss, validloss = [], []
        saver = tf.train.Saver()
        n = vdata.shape[1]
        ntrain = self.ntrain
        nvalid = n - ntrain
        self.log("%s %s" % (ntrain, nvalid))
        # :148 (the reference's pixel branch names an undefined `featreshape`; the intended tensor is the target frame)
        nn_err = tf.NNErr(test, nlen)
        validdata = vdata[:, ntrain:]
        traindata = vdata[:, :ntrain]
        rng = np.random.RandomState(self.seed)
        for itr in range(1, self.nitr):
            choicesrc = rng.choice(ntrain, batch_size)
            choicetgt = rng.choice(ntrain, batch_size)
            srcdata = traindata[np.arange(0, batch_size) % nlen, choicesrc]
            tgtdata = traindata[np.arange(0, batch_size) % nlen, choicetgt]
            tgtctx = traindata[0, choicetgt]
            batch = [srcdata, tgtctx, tgtdata]
            if itr % 4 == 0:
                _, loss, sim, r1, r2, err = sess.run([optimizer, test.loss, test.simloss, test.recon1, test.recon2,
                                                      nn_err], {tfinput: batch, learning_rate: 1e-4})
                self.log("%s %s %s %s %s %s" % (itr, loss, sim, r1, r2, err))
                allloss.append(loss)
                self.history.append(dict(Iteration=itr, Loss=loss, Sim=sim, R1=r1, R2=r2, NNErr=err))
            else:
                sess.run(optimizer, {tfinput: batch, learning_rate: 1e-4})

            if (itr % 40 == 0 or itr % self.save_every == 0) and nvalid > 0:
                choicesrc = rng.choice(nvalid, batch_size)
                choicetgt = rng.choice(nvalid, batch_size)
                srcdata = validdata[np.arange(0, batch_size) % nlen, choicesrc]
                tgtdata = validdata[np.arange(0, batch_size) % nlen, choicetgt]
                tgtctx = validdata[0, choicetgt]
                batch = [srcdata, tgtctx, tgtdata]
                loss, sim, r1, r2, err = sess.run([test.loss, test.simloss, test.recon1, test.recon2, nn_err],
                                                  {tfinput: batch})
                self.log("%s %s %s %s %s %s E" % (itr, loss, sim, r1, r2, err))
                validloss.append(loss)
                if itr % self.save_every == 0:
                    # train_script.py:187-190 formats 'model/<itr>/model_...' with a basedir that ends in '/': one
                    # normalised directory for both files, whatever the caller passed
                    ckdir = os.path.join(self.basedir, str(itr))
                    os.makedirs(ckdir, exist_ok=True)
                    saver.save(sess, os.path.join(ckdir, "model_%d_%.2f_%.2f_%.2f_%d" % (itr, loss, r1, r2, err)))
                    np.save(os.path.join(ckdir, "validloss"), validloss)
        self.sess, self.test = sess, test
        return allloss, validloss
