"""Reward featurizer: the model-dependent part of ``BaseSampler.initialize`` / ``process_samples``
(rllab/sampler/base.py:57-156, 166-257), batched for B200.

The reference issues one ``sess.run`` per trajectory per viewpoint with B=25 (250 tiny launches per TRPO
iteration).  Here all trajectories of an iteration are stacked into one uint8 batch, copied to
the device once, pushed through the source encoder in one launch per layer, and reduced to rewards by one
distance kernel that writes straight into the float64 reward array:

    cost[p,j]   = sum_k (means[vp][j,k] - z[p,j,k])^2 + scale * sum_hwc (imgs[vp][j] - x[p,j])^2   (base.py:243-245)
    rewards[p][2j+1] -= cost[p,j] * j^2                                                            (base.py:256-257)

Everything downstream of the rewards (GAE, baselines, logging; base.py:259-395) is generic rllab bookkeeping
and stays in the caller.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops
from .models import ContextAEReal, ContextSkipNew, _ContextModel

KWARG_KEYS = ("mode", "mode2", "scale", "nvp", "imsize", "name", "modelname", "modeldata", "meanfile", "layer",
              "experttheano", "ablation_type", "vp", "angle", "goal", "object", "geoms")     # SURVEY.md §5 "Config"


class RewardFeaturizer:
    """Holds the restored model plus the expert statistics; `process_paths` is the per-iteration hot call."""

    T = 25                                  # frames per trajectory: self.batch_size = 25 (base.py:115)

    def __init__(self, mode: str = "ours", name: str = "push", imsize=(48, 48), nvp: int = 1, scale: float = 1.0,
                 modelname: Optional[str] = None, modeldata: Optional[str] = None, ablation_type: str = "None",
                 precision: str = "mixed", model: Optional[_ContextModel] = None, device="cuda", layer: str = "Mixed_7c",
                 meanfile: Optional[str] = None, inception=None, chunk_paths: int = 8, **unused_kwargs):
        if not (mode.startswith("ours") or mode.startswith("inception")):
            raise _lib.CtxError("RewardFeaturizer covers mode 'ours*' (base.py:192-257) and 'inception*' (:69-111,179-189)")
        self.layer, self.meanfile, self.inception, self.chunk_paths = layer, meanfile, inception, int(chunk_paths)
        self.std: Optional[torch.Tensor] = None
        if ablation_type not in ("None", "nofeat", "noimage"):
            raise _lib.CtxError("ablation_type %r is broken in the reference (base.py:250-252 uses an undefined "
                                "image_recon) and is not provided" % ablation_type)
        self.mode, self.name, self.imsize, self.nvp, self.scale = mode, name, tuple(imsize), int(nvp), float(scale)
        self.modelname, self.modeldata, self.ablation_type = modelname, modeldata, ablation_type
        self.precision, self.device = precision, device
        self.model = model
        self.means: List[torch.Tensor] = []
        self.imgs: List[torch.Tensor] = []
        self.initialized = False

    # -- base.py:113-154 ---------------------------------------------------------------------------------------------
    def initialize(self):
        if self.mode.startswith("inception"):
            # base.py:69-111: InceptionV3(is_training=False); expert statistics from `meanfile` (layer, layer+'std')
            from .inception import InceptionV3
            if self.inception is None:
                self.inception = InceptionV3(device=self.device, precision=self.precision)
                if self.modelname:                                # base.py:108-110: slim's inception_v3.ckpt (or an .npz)
                    from . import tf_checkpoint
                    self.inception.load_state_dict(tf_checkpoint.load_variables(self.modelname, list(self.inception.shapes))[0])
            if self.meanfile:
                data = np.load(self.meanfile)
                self.set_inception_statistics(data[self.layer], data[self.layer + "std"])
            self.initialized = True
            return
        if self.mode == "oursinception" and self.model is None:
            # base.py:121-132: frozen InceptionV3 -> Mixed_7c feature maps [3,25,8,8,2048] -> ContextAEInception2
            from .inception import InceptionV3
            from .models import ContextAEInception2
            if self.inception is None:
                self.inception = InceptionV3(device=self.device, precision=self.precision)
            self.model = ContextAEInception2(strides=[1, 2, 1, 2], kernels=[3, 3, 3, 3], filters=[1024, 1024, 512, 512])
            fh = self._inception_map_size(self.imsize[0])
            fw = self._inception_map_size(self.imsize[1])
            self.model.build(torch.empty((3, self.T, fh, fw, 2048), dtype=torch.float32, device="meta"),
                             device=self.device, precision=self.precision)
            if self.modelname:
                # base.py:143-145 restores every non-optimiser variable of the checkpoint: the context model and, when the
                # file holds them (train_script.py saves the whole graph), the frozen InceptionV3 variables
                from . import compat, tf_checkpoint
                compat.restore_model(self.model.net, self.modelname)
                try:
                    inc_sd, _ = tf_checkpoint.load_variables(self.modelname, list(self.inception.shapes))
                    self.inception.load_state_dict(inc_sd)
                except KeyError:
                    pass                                   # a context-model-only file: InceptionV3 is loaded separately
        if self.model is None:
            if self.name in ("real", "sweep"):
                self.model = ContextAEReal()                                   # base.py:134-135
            elif self.name in ("push", "reach", "strike", "throw"):
                self.model = ContextSkipNew()                                  # base.py:136-137
            else:
                raise _lib.CtxError("unknown task name %r" % self.name)
            shape = (3, self.T) + self.imsize + (3,)
            self.model.build(torch.empty(shape, dtype=torch.uint8, device="meta"), device=self.device,
                             precision=self.precision)
            if self.modelname:
                from . import compat
                compat.restore_model(self.model.net, self.modelname)          # base.py:143-145
        self.initialized = True

    @staticmethod
    def _inception_map_size(n: int) -> int:
        """Spatial size of Mixed_7c for an n-pixel side (VALID/stride chain of nets/inception_v3.py:100-378)."""
        n = (n - 3) // 2 + 1          # Conv2d_1a_3x3 s2 VALID
        n = n - 2                     # Conv2d_2a_3x3 VALID   (2b is SAME)
        n = (n - 3) // 2 + 1          # MaxPool_3a
        n = n - 2                     # Conv2d_4a_3x3 VALID   (3b is 1x1)
        n = (n - 3) // 2 + 1          # MaxPool_5a
        n = (n - 3) // 2 + 1          # Mixed_6a
        return (n - 3) // 2 + 1       # Mixed_7a

    def _model_input(self, frames_u8: torch.Tensor) -> torch.Tensor:
        """What the context model sees for uint8 frames [n,H,W,3]: the normalised frames (base.py:116-119), or — mode
        'oursinception' — their Mixed_7c feature maps (base.py:122-131; `image_trans` is then the feature tensor)."""
        if self.mode == "oursinception":
            n, chunk = frames_u8.shape[0], 128                       # bound the InceptionV3 activation footprint
            if n <= chunk:
                return self.inception.forward(frames_u8, ("Mixed_7c",))["Mixed_7c"].clone()
            out = None
            for c0 in range(0, n, chunk):
                f = self.inception.forward(frames_u8[c0:c0 + chunk], ("Mixed_7c",))["Mixed_7c"]
                if out is None:
                    out = torch.empty((n,) + tuple(f.shape[1:]), device=f.device, dtype=torch.float32)
                out[c0:c0 + f.shape[0]].copy_(f)
            return out
        return ops.u8_normalize(frames_u8.contiguous())

    # -- base.py:195-223: expert statistics, once per process ----------------------------------------------------------
    def compute_expert_means(self, validdata: np.ndarray, contexts: Sequence[np.ndarray], chunk_videos: int = 64):
        """validdata [nlen, N, H, W, 3] float in [-1,1] (the trainer's vdata .npy); contexts[vp] = first frame of the
        agent's first trajectory (uint8 [H,W,3]).  Sets self.means[vp] [T,F] and self.imgs[vp] [T,H,W,3]."""
        if not self.initialized:
            self.initialize()
        skip = 2 if self.name in ("real", "sweep") else 1                      # base.py:208-210
        if self.mode == "oursinception":
            demo = np.asarray(validdata[::skip, :50]).astype(np.uint8)        # base.py:204-205,211-212: 50 videos, as stored
        else:
            demo = ((validdata[::skip] + 1) * 127.5).astype(np.uint8)         # base.py:214, [T, N, H, W, 3]
        if demo.shape[0] != self.T:
            raise _lib.CtxError("demos have %d frames after skip, expected %d" % (demo.shape[0], self.T))
        N = demo.shape[1]
        m = self.model
        self.means, self.imgs = [], []
        for vp in range(self.nvp):
            ctx = torch.as_tensor(np.ascontiguousarray(contexts[vp])).to(self.device)[None]   # [1,H,W,3] u8
            ctx = self._model_input(ctx)                                                        # frames or feature maps
            z_ctx, skips = m.enc2(ctx)
            per_frame = tuple(ctx.shape[1:])
            sum_z = torch.zeros((self.T, m.net.F), device=self.device, dtype=torch.float64)
            sum_img = torch.zeros((self.T,) + per_frame, device=self.device, dtype=torch.float64)
            for v0 in range(0, N, chunk_videos):
                vids = demo[:, v0:v0 + chunk_videos]                             # [T, n, H, W, 3]
                n = vids.shape[1]
                frames = torch.as_tensor(np.ascontiguousarray(vids.transpose(1, 0, 2, 3, 4))).to(self.device)
                frames = self._model_input(frames.reshape((n * self.T,) + self.imsize + (3,)))
                z_src = m.enc1(frames)
                tz = m.translate(z_src, z_ctx.expand(n * self.T, -1))
                out = m.dec(tz, skips, ctx_frames=ctx)
                sum_z += tz.view(n, self.T, -1).sum(0, dtype=torch.float64)
                sum_img += out.view((n, self.T) + per_frame).sum(0, dtype=torch.float64)
            self.means.append((sum_z / N).float().contiguous())
            self.imgs.append((sum_img / N).float().contiguous())

    # -- base.py:94-110: expert statistics of the Inception baseline ------------------------------------------------
    def set_inception_statistics(self, means: np.ndarray, std: np.ndarray):
        """means / std [T, h, w, c] of the chosen end point over expert rollouts (np.mean / np.std, base.py:103-104)."""
        self.means = [torch.as_tensor(np.ascontiguousarray(means), dtype=torch.float32).to(self.device)]
        self.std = torch.as_tensor(np.ascontiguousarray(std), dtype=torch.float32).to(self.device)

    def inception_features(self, frames_u8: torch.Tensor) -> torch.Tensor:
        """End point `layer` for uint8 frames [n, H, W, 3] on the device (base.py:183)."""
        return self.inception.forward(frames_u8, (self.layer,))[self.layer]

    def _process_paths_inception(self, paths: List[dict]) -> np.ndarray:
        """base.py:179-189, batched: `chunk_paths` trajectories (x 25 frames) per InceptionV3 pass, one reward kernel."""
        if not self.means or self.std is None:
            raise _lib.CtxError("Inception statistics missing: pass meanfile= or call set_inception_statistics()")
        P, H, W = len(paths), self.imsize[0], self.imsize[1]
        costs = np.zeros((P, self.T), np.float32)
        for c0 in range(0, P, self.chunk_paths):
            chunk = paths[c0:c0 + self.chunk_paths]
            n = len(chunk)
            L = max(len(p["rewards"]) for p in chunk)
            frames = torch.empty((n, self.T, H, W, 3), dtype=torch.uint8).pin_memory()
            rewards = torch.zeros((n, L), dtype=torch.float64).pin_memory()
            for i, path in enumerate(chunk):
                imgs = [im[0] for im in path["env_infos"]["imgs"] if im is not None]          # base.py:182
                if len(imgs) != self.T:
                    raise _lib.CtxError("path %d has %d frames, expected %d" % (c0 + i, len(imgs), self.T))
                frames.numpy()[i] = np.asarray(imgs)
                rewards.numpy()[i, :len(path["rewards"])] = path["rewards"]
            feat = self.inception_features(frames.to(self.device, non_blocking=True).view(n * self.T, H, W, 3))
            d_rewards = rewards.to(self.device, non_blocking=True)
            d_cost = torch.zeros((n, self.T), device=self.device, dtype=torch.float32)
            ops.reward_inception(feat, self.means[0], self.std, n, self.T, rewards=d_rewards, cost=d_cost)
            out = d_rewards.cpu().numpy()
            costs[c0:c0 + n] = d_cost.cpu().numpy()
            for i, path in enumerate(chunk):
                path["rewards"][:] = out[i, :len(path["rewards"])]
        return costs

    # -- base.py:232-257: the hot loop ---------------------------------------------------------------------------------
    def process_paths(self, paths: List[dict]) -> np.ndarray:
        """Subtract the feature/pixel distance from every path's rewards, in place; returns the costs [paths,T]."""
        if not self.initialized:
            self.initialize()
        if self.mode.startswith("inception"):
            return self._process_paths_inception(paths)
        if not self.means:
            raise _lib.CtxError("expert statistics missing: call compute_expert_means() first (base.py:195-223)")
        P = len(paths)
        if P == 0:
            return np.zeros((0, self.T), np.float32)
        H, W = self.imsize
        L = max(len(p["rewards"]) for p in paths)
        if L < 2 * self.T:
            raise _lib.CtxError("rewards must have at least %d entries (index 2j+1, j<25)" % (2 * self.T))
        # gather host side: frames [nvp, P, T, H, W, 3] uint8 and rewards [P, L] float64, pinned for async copies
        frames = torch.empty((self.nvp, P, self.T, H, W, 3), dtype=torch.uint8).pin_memory()
        rewards = torch.zeros((P, L), dtype=torch.float64).pin_memory()
        fr_np, rw_np = frames.numpy(), rewards.numpy()
        for i, path in enumerate(paths):
            imgs = [im for im in path["env_infos"]["imgs"] if im is not None]          # base.py:193
            if len(imgs) != self.T:
                raise _lib.CtxError("path %d has %d frames, expected %d" % (i, len(imgs), self.T))
            arr = np.asarray(imgs)                                                       # [T, nvp, H, W, 3]
            for vp in range(self.nvp):
                fr_np[vp, i] = arr[:, vp]
            rw_np[i, :len(path["rewards"])] = path["rewards"]
        d_frames = frames.to(self.device, non_blocking=True)
        d_rewards = rewards.to(self.device, non_blocking=True)
        d_cost = torch.zeros((P, self.T), device=self.device, dtype=torch.float32)
        fw = 0.0 if self.ablation_type == "nofeat" else 1.0
        iw = 0.0 if self.ablation_type == "noimage" else 1.0
        total = torch.zeros((P, self.T), device=self.device, dtype=torch.float32)
        for vp in range(self.nvp):
            u8 = d_frames[vp].reshape((P * self.T, H, W, 3))
            x = self._model_input(u8)                                                    # image_trans[0]
            z = self.model.enc1(x)                                                       # input_z, base.py:234
            ops.reward_ours(z, x, self.means[vp], self.imgs[vp], P, self.T, self.scale, rewards=d_rewards,
                            cost=d_cost, feat_weight=fw, img_weight=iw)
            total += d_cost
        out_rewards = d_rewards.cpu().numpy()
        for i, path in enumerate(paths):
            path["rewards"][:] = out_rewards[i, :len(path["rewards"])]
        return total.cpu().numpy()
