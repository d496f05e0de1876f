"""Deterministic synthetic demo-video bank (SURVEY.md §8d "Synthetic inputs").

The reference trains on gym-monitor mp4 demos decoded to ``vdata[nlen, N, H, W, 3]`` (scripts/train_script.py:59-96);
neither the demos nor the simulator ship, so every test and benchmark uses this generator instead: per video a
random background colour, a random-colour 12x12 square (scaled with the frame) moving linearly from a random
start to a random goal over ``nlen`` frames, plus uniform noise in [-4, 4].  uint8, layout [nlen, N, H, W, 3].
"""
from __future__ import annotations

import numpy as np


def video_bank(nlen: int = 25, nvideos: int = 64, height: int = 64, width: int = 64, channels: int = 3,
               seed: int = 1234) -> np.ndarray:
    rng = np.random.RandomState(seed)
    bank = np.empty((nlen, nvideos, height, width, channels), dtype=np.uint8)
    side = max(2, (12 * min(height, width)) // 64)
    for v in range(nvideos):
        bg = rng.randint(32, 224, size=channels)
        fg = rng.randint(0, 256, size=channels)
        start = rng.uniform(0, 1, size=2)
        goal = rng.uniform(0, 1, size=2)
        for t in range(nlen):
            a = t / max(nlen - 1, 1)
            cy = int(round(((1 - a) * start[0] + a * goal[0]) * (height - side)))
            cx = int(round(((1 - a) * start[1] + a * goal[1]) * (width - side)))
            frame = np.empty((height, width, channels), dtype=np.int16)
            frame[:] = bg
            frame[cy:cy + side, cx:cx + side] = fg
            frame += rng.randint(-4, 5, size=frame.shape).astype(np.int16)
            bank[t, v] = np.clip(frame, 0, 255).astype(np.uint8)
    return bank


def to_float(bank_u8: np.ndarray) -> np.ndarray:
    """The trainer's float path: ``x/127.5 - 1`` (scripts/train_script.py:16-20 ``transform`` with rescale)."""
    return bank_u8.astype(np.float64) / 127.5 - 1.0


def trainer_indices(batch_size: int, nlen: int, ntrain: int, rng: np.random.RandomState):
    """scripts/train_script.py:154-158: (time index per element, src video ids, tgt video ids)."""
    choicesrc = rng.choice(ntrain, batch_size)
    choicetgt = rng.choice(ntrain, batch_size)
    t = np.arange(0, batch_size) % nlen
    return t, choicesrc, choicetgt


def trainer_batch(bank: np.ndarray, batch_size: int, nlen: int, ntrain: int, rng: np.random.RandomState) -> np.ndarray:
    """[3, B, H, W, C] = [src frames, tgt first frames (context), tgt frames]  (train_script.py:156-159)."""
    t, cs, ct = trainer_indices(batch_size, nlen, ntrain, rng)
    return np.stack([bank[t, cs], bank[0, ct], bank[t, ct]])
