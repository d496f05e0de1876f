"""Host-side frame preparation of the reference (data prep only — not on the device hot path).

    transform / inverse_transform   gym/envs/mujoco/arm_shaping.py:12-16, scripts/train_script.py:16-22
    demo bank layout                scripts/train_script.py:83-97: vdata [nlen, N, H, W, 3]

`scipy.misc.imresize` (removed from SciPy long ago) was a PIL resize with bilinear interpolation of the uint8 image;
it is restated here with Pillow.  The mp4 / ffmpeg front-end of the trainer (train_script.py:59-82) is out of scope:
ModelTrainer takes the `.npy` bank the reference saves.
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
