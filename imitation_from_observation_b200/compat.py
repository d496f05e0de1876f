"""`tf`-shaped adapter so the reference's call sites keep working with `import ...compat as tf`.

Covers exactly what scripts/train_script.py:99-203 and rllab/sampler/base.py:113-154,216,234 use on this path:

    placeholder, uint8/float32, image.convert_image_dtype, subtract, multiply, variable_scope,
    Session().run(fetches, feed_dict), train.AdamOptimizer(lr).minimize(loss, var_list=...),
    train.Saver().save/restore, global_variables_initializer, get_collection, GraphKeys, ConfigProto.

`Session.run` does what TF's executor did: find the smallest forward level that produces every fetch
(`input_z` alone runs only the source encoder — the pruning base.py:234 relies on), copy the fed batch host ->
device through pinned memory, launch, and copy the fetched values back as numpy arrays.
"""
from __future__ import annotations

import contextlib
import os
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .models import Fetch, _ContextModel
from .network import FWD_ALL

uint8 = torch.uint8
float32 = torch.float32

_scope_stack: List[str] = []


@contextlib.contextmanager
def variable_scope(name: str):
    _scope_stack.append(name)
    try:
        yield name
    finally:
        _scope_stack.pop()


def current_scope() -> str:
    return "".join(s + "/" for s in _scope_stack)


class Placeholder:
    """tf.placeholder: a typed shape plus the (only) preprocessing chain the reference applies to it."""

    def __init__(self, dtype, shape, name=None, chain=(), root=None):
        self.dtype, self.shape, self.name = dtype, tuple(int(s) for s in shape), name
        self.chain = tuple(chain)
        self.root = root or self
        if root is None:
            self.consumers: List[_ContextModel] = []

    def _derive(self, op):
        return Placeholder(self.dtype, self.shape, self.name, self.chain + (op,), self.root)

    def _register_consumer(self, model):
        norm = (("convert_image_dtype",), ("subtract", 0.5), ("multiply", 2.0))
        if self.root.dtype == torch.uint8:
            chain = self.chain[:-1] if self.chain and self.chain[-1][0] == "inception" else self.chain
            if chain != norm:
                raise _lib.CtxError("a uint8 placeholder must be normalised exactly as base.py:117-119 "
                                    "(convert_image_dtype, subtract 0.5, multiply 2.0); got %r" % (self.chain,))
        elif self.chain:
            raise _lib.CtxError("float placeholders are consumed as-is (train_script.py:118)")
        self.root.consumers.append(model)


def placeholder(dtype, shape=(), name=None) -> Placeholder:
    return Placeholder(dtype, shape, name)


class image:                                                      # noqa: N801  (mirrors tf.image)
    @staticmethod
    def convert_image_dtype(x: Placeholder, dtype=float32) -> Placeholder:
        return x._derive(("convert_image_dtype",))


def inception_mixed_7c(x: Placeholder, net) -> Placeholder:
    """The frozen InceptionV3 front-end of train_script.py:100-115 / base.py:121-131: normalised frames [3,B,H,W,3]
    -> `Mixed_7c` feature maps reshaped to [3,B,h,w,2048] (what `test.build(featreshape)` consumes).  `net` is an
    `inception.InceptionV3`; it is evaluated by Session.run on every feed, never trained."""
    from .sampler import RewardFeaturizer
    three, B, H, W, _ = x.shape
    fh, fw = RewardFeaturizer._inception_map_size(H), RewardFeaturizer._inception_map_size(W)
    out = Placeholder(x.dtype, (three, B, fh, fw, 2048), x.name, x.chain + (("inception", net),), x.root)
    return out


def subtract(x: Placeholder, v: float) -> Placeholder:
    return x._derive(("subtract", float(v)))


def multiply(x: Placeholder, v: float) -> Placeholder:
    return x._derive(("multiply", float(v)))


class GraphKeys:
    TRAINABLE_VARIABLES = "trainable_variables"


_models: List[_ContextModel] = []


def get_collection(key, scope: str = "") -> List[str]:
    out = []
    for m in _models:
        out += [v for v in m.trainable_variables() if v.startswith(scope)]
    return out


def global_variables_initializer():
    return None


class ConfigProto:
    def __init__(self, **kw):
        class _G:
            allow_growth = True
        self.gpu_options = _G()


class _Minimize:
    def __init__(self, opt: "_Adam", loss: Fetch):
        self.opt, self.loss = opt, loss


class _Adam:
    def __init__(self, learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon

    def minimize(self, loss: Fetch, var_list=None) -> _Minimize:
        if not isinstance(loss, Fetch) or loss.name != "loss":
            raise _lib.CtxError("minimize() supports the model's .loss (train_script.py:128)")
        return _Minimize(self, loss)


class _Saver:
    """tf.train.Saver stand-in.  restore() reads a TensorFlow checkpoint — the tensor bundle `<path>.index` +
    `<path>.data-*` a reference-trained model is stored in (tf_checkpoint.py) — or, failing that, the `.npz` keyed by the
    same TF variable names that save() writes by default.  `Saver(format="tf")` (or CTX_SAVER_FORMAT=tf)
    makes save() write the TensorFlow format instead, so the reference's TF code can restore models trained here."""

    def __init__(self, var_list=None, format: Optional[str] = None):
        self.format = format or os.environ.get("CTX_SAVER_FORMAT", "npz")
        if self.format not in ("npz", "tf"):
            raise _lib.CtxError("Saver format must be 'npz' or 'tf'")

    @staticmethod
    def _path(p):
        return p if p.endswith(".npz") else p + ".npz"

    def save(self, sess, path):
        sd = {}
        for m in _models:
            for k, v in m.net.state_dict().items():
                sd[m.scope + k] = v
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        if self.format == "tf":
            from . import tf_checkpoint
            tf_checkpoint.write_checkpoint(path, sd)
        else:
            np.savez(self._path(path), **sd)
        return path

    def restore(self, sess, path):
        for m in _models:
            restore_model(m.net, path, scope=m.scope)


def restore_model(net, path: str, scope: str = ""):
    """Load `net`'s variables from a TensorFlow checkpoint (V2 bundle or V1 file) or the .npz Saver.save writes, trying
    the model's own scope, then "contextmodel/", then no prefix (tf_checkpoint.load_variables).  Shared by Saver.restore
    and sampler.RewardFeaturizer.initialize (base.py:143-145)."""
    from . import tf_checkpoint
    sd, pre = tf_checkpoint.load_variables(path, list(net.shapes), prefixes=(scope, "contextmodel/", ""))
    net.load_state_dict(sd, prefix=pre)


class train:                                                      # noqa: N801  (mirrors tf.train)
    AdamOptimizer = _Adam
    Saver = _Saver


def register_model(model: _ContextModel):
    """Called by build_model(); makes the model visible to Saver / get_collection."""
    if model not in _models:
        _models.append(model)


def reset_default_graph():
    _models.clear()


def build_model(model: _ContextModel, image: Placeholder, **kw):
    """`with tf.variable_scope("contextmodel"): model.build(image)` (train_script.py:119-121) in one call."""
    model.build(image, scope=current_scope(), **kw)
    register_model(model)
    return model


class Session:
    def __init__(self, config=None):
        self._pinned: Dict[Any, torch.Tensor] = {}
        self._pin_event: Dict[int, Any] = {}          # staging buffer -> event of the last H2D copy that read it
        self._prefetch_of: Dict[int, Any] = {}        # model -> (fed object, staged tensor) of a pending prefetch

    def _stage(self, ph: Placeholder, value) -> torch.Tensor:
        """Host value -> pinned staging buffer (-> device for uint8 frames), as one [3,B,H,W,C] tensor.  A float batch
        stays in pinned HOST memory: ContextNet.set_input copies it role by role, overlapped with the first kernels."""
        if torch.is_tensor(value) and value.is_cuda:
            return value
        if isinstance(value, (list, tuple)) and len(value) == 3 and all(torch.is_tensor(v) and v.is_cuda for v in value):
            # [src, ctx, tgt] device batches (data.DeviceBank): no staging, no host copy
            return torch.stack(list(value)) if ph.chain else list(value)
        want_t = torch.uint8 if ph.root.dtype == torch.uint8 else torch.float32
        host_ok = want_t == torch.float32 and not ph.chain
        if torch.is_tensor(value) and value.dtype == want_t and value.is_contiguous() and value.is_pinned():
            return value if host_ok else value.to("cuda", non_blocking=True)   # caller already staged it in pinned memory
        arr = np.asarray(value)
        want = np.uint8 if ph.root.dtype == torch.uint8 else np.float32
        key = (id(ph.root), arr.shape, want)
        pin = self._pinned.get(key)
        if pin is None:
            pin = torch.empty(arr.shape, dtype=torch.uint8 if want is np.uint8 else torch.float32).pin_memory()
            self._pinned[key] = pin
        ev = self._pin_event.get(pin.data_ptr())
        if ev is not None:
            ev.synchronize()                                     # the previous step's async copy out of this buffer
        np.copyto(pin.numpy(), arr, casting="unsafe")            # f64 -> f32 cast happens here (train_script feed)
        if host_ok:
            return pin
        # uint8 frames (and chained float feeds) are copied to the device here: remember the copy, so that the next
        # _stage() does not refill the pinned buffer while this H2D copy is still queued behind earlier kernels
        # (minimize-only steps never synchronise the host)
        dev = pin.to("cuda", non_blocking=True)
        self._pin_event[pin.data_ptr()] = torch.cuda.current_stream().record_event()
        return dev

    def prefetch(self, feed_dict: dict):
        """Start the host->device copy of the NEXT step's batch while the current step computes (what tf.data's
        prefetch_to_device does for a TF input pipeline).  The next run() fed with the same object only flips the
        network's input slots.  Float image placeholders without a preprocessing chain only; other feeds are ignored."""
        for k, v in feed_dict.items():
            if not isinstance(k, Placeholder) or k.shape == () or k.chain or k.root.dtype == torch.uint8:
                continue
            for m in k.root.consumers:
                staged = self._stage(k, v)
                if torch.is_tensor(staged) and not staged.is_cuda:
                    self._pin_event[staged.data_ptr()] = m.net.prefetch_input(staged)
                    self._prefetch_of[id(m)] = (v, staged)

    def run(self, fetches, feed_dict: Optional[dict] = None, prefetch: Optional[dict] = None):
        """`prefetch`: the next call's feed_dict — its image batch is copied to the device under this step's kernels."""
        if fetches is None:                                   # sess.run(tf.global_variables_initializer())
            return None
        single = not isinstance(fetches, (list, tuple))
        flist = [fetches] if single else list(fetches)
        # a normalised-image placeholder may itself be fetched (base.py:234 fetches self.image_trans)
        flist = [Fetch(f.root.consumers[0], "image_trans") if isinstance(f, Placeholder) and f.root.consumers else f
                 for f in flist]
        feed_dict = feed_dict or {}
        models: List[_ContextModel] = []
        for f in flist:
            m = f.loss.model if isinstance(f, _Minimize) else getattr(f, "model", None)
            if m is not None and m not in models:
                models.append(m)
        lr_value = None
        for k, v in feed_dict.items():
            if isinstance(k, Placeholder) and k.shape == ():
                lr_value = float(v)
        results: Dict[int, Any] = {}
        for m in models:
            register_model(m)
            root = m.placeholder.root if isinstance(m.placeholder, Placeholder) else None
            if root is not None:
                fed = [v for k, v in feed_dict.items() if isinstance(k, Placeholder) and k.root is root]
                if not fed:
                    raise _lib.CtxError("no value fed for the model's image placeholder")
                pre = self._prefetch_of.pop(id(m), None)
                hit = pre is not None and pre[0] is fed[0]      # the batch a prefetch() already put in flight
                staged = pre[1] if hit else self._stage(m.placeholder, fed[0])
                chain = m.placeholder.chain
                if chain and chain[-1][0] == "inception":      # frozen front-end: frames -> Mixed_7c feature maps
                    inc = chain[-1][1]
                    three, Bf = staged.shape[0], staged.shape[1]
                    feats = inc.forward(staged.reshape((three * Bf,) + tuple(staged.shape[2:])), ("Mixed_7c",))["Mixed_7c"]
                    staged = feats.view((three, Bf) + tuple(feats.shape[1:]))
                m.net.set_input(staged, use_prefetched=hit)
                if torch.is_tensor(staged) and not staged.is_cuda and m.net._in_events is not None:
                    self._pin_event[staged.data_ptr()] = m.net._in_events[1]
            mine = [f for f in flist if (f.loss.model if isinstance(f, _Minimize) else getattr(f, "model", None)) is m]
            mins = [f for f in mine if isinstance(f, _Minimize)]
            level = max([FWD_ALL] if mins else [f.level for f in mine])
            if mins:
                opt = mins[0].opt
                lr = opt.lr
                if isinstance(lr, Placeholder):
                    if lr_value is None:
                        raise _lib.CtxError("learning-rate placeholder was not fed")
                    lr = lr_value
                graphed = m.net.step_graphed(float(lr), opt.b1, opt.b2, opt.eps)
                if not graphed:
                    m.net.zero_grad()
                    m.net.forward(FWD_ALL)
                # fetched values must be those of the pre-update graph, like TF.  Backward and Adam never write a forward
                # buffer, so after a whole-step graph replay the activations and losses still are the pre-update ones.
                for f in mine:
                    if isinstance(f, Fetch):
                        results[id(f)] = m.net.fetch(f.name).detach().clone()
                    elif isinstance(f, NNErr):
                        results[id(f)] = m.net.nn_err(f.nlen)
                if not graphed:
                    m.net.apply_gradients(float(lr), opt.b1, opt.b2, opt.eps)
            else:
                m.net.forward(max(level, 1))
                for f in mine:
                    results[id(f)] = m.net.nn_err(f.nlen) if isinstance(f, NNErr) else m.net.fetch(f.name)
        if prefetch:
            self.prefetch(prefetch)                    # enqueued behind this step's launches, ahead of the blocking reads
        out = []
        for f in flist:
            if isinstance(f, _Minimize):
                out.append(None)
            elif isinstance(f, NNErr):
                out.append(results[id(f)])
            else:
                v = results[id(f)].detach().cpu().numpy()
                out.append(v.item() if v.ndim == 0 else v)
        return out[0] if single else out

    def close(self):
        self._pinned.clear()


class NNErr:
    """Handle for the trainer's nearest-neighbour frame-index error (train_script.py:148)."""

    def __init__(self, model: _ContextModel, nlen: int):
        self.model, self.nlen, self.level = model, int(nlen), FWD_ALL
        self.name = "nn_err"
