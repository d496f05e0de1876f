"""Host-side frame preparation of the reference (data prep only — not on the device hot path).

    transform / inverse_transform   gym/envs/mujoco/arm_shaping.py:12-16, scripts/train_script.py:16-22
    demo bank layout                scripts/train_script.py:83-97: vdata [nlen, N, H, W, 3]

`scipy.misc.imresize` (removed from SciPy long ago) was a PIL resize with bilinear interpolation of the uint8 image;
it is restated here with Pillow.  The mp4 front-end of the trainer (train_script.py:59-96, imageio + ffmpeg there) is
`read_video` / `videos_to_bank` on OpenCV's FFMPEG reader; ModelTrainer also takes the `.npy` bank the reference saves.

`DeviceBank` keeps that bank in HBM and answers the trainer's own indexing expressions (train_script.py:150-173) with
device batches assembled by `ctx_gather_frames`, which removes the per-step host gather, float64->float32 cast and H2D
copy.
"""
from __future__ import annotations

import numpy as np


def imresize(image: np.ndarray, size) -> np.ndarray:
    """scipy.misc.imresize(image, [height, width]) for uint8 HxWxC images: PIL bilinear, uint8 out."""
    from PIL import Image
    h, w = int(size[0]), int(size[1])
    img = Image.fromarray(np.asarray(image, dtype=np.uint8))
    return np.asarray(img.resize((w, h), resample=Image.BILINEAR))


def transform(image: np.ndarray, resize_height: int, resize_width: int, rescale: bool = True) -> np.ndarray:
    """train_script.py:16-20 / arm_shaping.py:12-14: resize, then x / 127.5 - 1 (float64) unless rescale is False."""
    cropped = imresize(image, [resize_height, resize_width])
    if rescale:
        return np.array(cropped) / 127.5 - 1.0
    return cropped


def inverse_transform(images: np.ndarray) -> np.ndarray:
    """train_script.py:21-22 / arm_shaping.py:15-16: (x + 1) / 2."""
    return (images + 1.0) / 2.0


def frames_to_bank(videos, nlen: int, nskip: int, idims, rescale: bool = True) -> np.ndarray:
    """train_script.py:67-97 without the video decoder: `videos` yields sequences of 51 decoded frames; frames
    1, 1+nskip, ... < 51 are transformed and stacked time-major -> vdata [nlen, N, H, W, 3]."""
    idata = [[] for _ in range(nlen)]
    for vid in videos:
        if len(vid) != 51:                                     # train_script.py:72
            continue
        frames = [transform(vid[j], idims[0], idims[1], rescale) for j in range(1, 51, nskip)]
        if rescale and any(np.max(f) == -1 for f in frames):   # an all-black frame drops the video (:76-79)
            continue
        if len(frames) != nlen:
            continue
        for j, f in enumerate(frames):
            idata[j].append(f)
    return np.array(idata)


def read_video(path: str):
    """All frames of a video file as RGB uint8 arrays — imageio.get_reader(name, 'ffmpeg') of train_script.py:66 (same
    FFMPEG decoder underneath; OpenCV hands frames back as BGR)."""
    import cv2
    cap = cv2.VideoCapture(path)
    if not cap.isOpened():
        raise IOError("cannot open video %s" % path)
    frames = []
    try:
        while True:
            ok, frame = cap.read()
            if not ok:
                break
            frames.append(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB))
    finally:
        cap.release()
    return frames


def videos_to_bank(paths, nlen: int, nskip: int, idims, rescale: bool = True, nvideos: int = 1 << 30,
                   inception: bool = False, rng=None, log=None) -> np.ndarray:
    """train_script.py:58-96: shuffle the demo videos, keep those with exactly 51 frames, take frames 1, 1 + nskip, ... < 51,
    transform each (resize, x / 127.5 - 1 unless `rescale` is off), drop a video with an all-black frame (unless
    `inception`: uint8 frames go to the feature extractor), stop after `nvideos` videos or more than 10 unreadable files.
    Returns vdata [nlen, N, H, W, 3] — the array the reference saves as vdata_strike<N>.npy."""
    paths = list(paths)
    (rng or np.random).shuffle(paths)                          # :63
    idata = [[] for _ in range(nlen)]
    nfail = itr = 0
    for name in paths:
        try:
            vid = read_video(name)
            if len(vid) == 51:                                 # :69
                frames = []
                for j in range(1, 51, nskip):
                    frame = transform(vid[j], idims[0], idims[1], rescale)
                    if not inception and np.max(frame) == -1:  # :73-76
                        if log:
                            log("rip %s %s" % (itr, name))
                        frames = []
                        break
                    frames.append(frame)
                if len(frames) != nlen:
                    continue                                   # (:78-79: such a video does not count towards nvideos)
                for j, f in enumerate(frames):
                    idata[j].append(f)
            elif log:
                log("%s %s" % (name, len(vid)))
            itr += 1
        except Exception as ex:                                # noqa: BLE001  (:84-89: log, give up after 10 failures)
            nfail += 1
            if log:
                log("Unexpected error: %r %s" % (ex, name))
            if nfail > 10:
                break
        if itr >= nvideos:
            break
    return np.array(idata)


def pixel_table() -> np.ndarray:
    """The 256 values transform() can produce, in the reference's arithmetic: float32(v / 127.5 - 1.0) computed in
    float64 (arm_shaping.py:12-14) and cast by the float32 feed (train_script.py:118)."""
    return (np.arange(256, dtype=np.float64) / 127.5 - 1.0).astype(np.float32)


class DeviceBank:
    """The demo bank `vdata [nlen, N, H, W, C]` (train_script.py:97) resident in HBM.

    Supports exactly what the trainer's loop does with the array (train_script.py:150-173):

        bank.shape                      (nlen, N, H, W, C)
        bank[:, :ntrain], bank[:, ntrain:]          views over a range of videos (traindata / validdata)
        bank[np.arange(B) % nlen, choice]           -> [B, H, W, C] float32 CUDA tensor (one frame per row)
        bank[0, choice]                             -> the videos' first frames (the context frames)

    A float bank whose every value is one of transform()'s 256 outputs — any bank made by the reference's pipeline — is
    stored as uint8 (a quarter of the memory and of the gather's read traffic) and mapped back through `pixel_table`;
    the batches are bit-identical to float32(vdata)[...] either way (tests/test_ops_gpu.py).  Anything else is stored
    as float32.  A uint8 array is taken as raw pixels and gathered through the same table (rescale=True semantics).
    """

    def __init__(self, vdata, device="cuda", _parent=None, _v0=0, _v1=None):
        import torch
        from . import _lib
        if _parent is not None:                                   # a video-range view
            self.__dict__.update(_parent.__dict__)
            self._v0, self._v1 = _v0, _v1
            return
        _lib.load()
        arr = np.asarray(vdata)
        if arr.ndim != 5:
            raise _lib.CtxError("DeviceBank expects vdata [nlen, N, H, W, C], got shape %s" % (arr.shape,))
        self.nlen, self.N = int(arr.shape[0]), int(arr.shape[1])
        self.frame_shape = tuple(int(d) for d in arr.shape[2:])
        self.device = torch.device(device)
        table = pixel_table()
        if arr.dtype == np.uint8:
            q = arr
        else:
            # chunked so that a multi-GB float64 bank is never duplicated on the host
            q = np.empty(arr.shape, np.uint8)
            ok = True
            for t in range(self.nlen):
                f32 = arr[t].astype(np.float32)
                qt = np.clip(np.rint((arr[t].astype(np.float64) + 1.0) * 127.5), 0, 255).astype(np.uint8)
                if not np.array_equal(table[qt], f32):
                    ok = False
                    break
                q[t] = qt
            if not ok:
                q = None
        if q is not None:
            self.store = torch.from_numpy(np.ascontiguousarray(q)).to(self.device).view((self.nlen * self.N,) + self.frame_shape)
            self.lut = torch.from_numpy(table).to(self.device)
        else:
            self.store = torch.empty((self.nlen * self.N,) + self.frame_shape, device=self.device, dtype=torch.float32)
            for t in range(self.nlen):
                self.store[t * self.N:(t + 1) * self.N].copy_(torch.from_numpy(arr[t].astype(np.float32)))
            self.lut = None
        self._v0, self._v1 = 0, self.N

    @property
    def shape(self):
        return (self.nlen, self._v1 - self._v0) + self.frame_shape

    @property
    def quantised(self) -> bool:
        return self.lut is not None

    def __len__(self):
        return self.nlen

    def __getitem__(self, key):
        import torch
        from . import ops
        if not (isinstance(key, tuple) and len(key) == 2):
            raise IndexError("DeviceBank supports bank[:, a:b] and bank[time_index, video_index]")
        t, v = key
        n = self._v1 - self._v0
        if isinstance(t, slice) and t == slice(None) and isinstance(v, slice):
            a, b, step = v.indices(n)
            if step != 1:
                raise IndexError("DeviceBank video ranges must be contiguous")
            return DeviceBank(None, _parent=self, _v0=self._v0 + a, _v1=self._v0 + max(a, b))
        v = np.asarray(v, dtype=np.int64)
        t = np.asarray(t, dtype=np.int64)
        if v.ndim != 1 or t.ndim > 1:
            raise IndexError("DeviceBank[time, video]: video must be a 1-D index array, time an int or a 1-D array")
        t = np.broadcast_to(t, v.shape)
        if v.size and (v.min() < -n or v.max() >= n):
            raise IndexError("video index out of range for %d videos" % n)
        if t.size and (t.min() < -self.nlen or t.max() >= self.nlen):
            raise IndexError("time index out of range for %d frames" % self.nlen)
        v = np.where(v < 0, v + n, v) + self._v0
        t = np.where(t < 0, t + self.nlen, t)
        flat = torch.from_numpy(np.ascontiguousarray(t * self.N + v)).to(self.device)
        return ops.gather_frames(self.store, flat, lut=self.lut)
