"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference batch augmentations (next row N3).

Follows ``SpecAugment.forward`` / ``specaugment`` (examples/speech_recognition/modules/specaugment.py:55-112)
and ``TimeStretch.forward`` / ``time_stretch_seq`` (modules/time_stretch.py:18-54) in numpy, with the
same calls to Python's ``random`` and ``numpy.random`` in the same order, so that under the same
seeds the results are the reference's.  ``torch.linspace`` (fp32) is restated from its published
arithmetic: step = (end - start) / (steps - 1); element i is start + step*i in the first half and
end - step*(steps - 1 - i) in the second half; ``torch.round`` is round-half-to-even.

Pinned by ``tests/golden/augment.pt`` (outputs of the live reference modules under fixed seeds,
generator ``oracle/make_golden_augment.py``) and, where /root/reference is mounted, by the live
modules directly (``tests/test_oracle_augment.py``).  Only tests/ may import this module.
"""
import random

import numpy as np


def specaugment_batch(x, frequency_masking_pars, time_masking_pars, frequency_masking_num,
                      time_masking_num, rate=1.0):
    """specaugment.py:55-67 over a padded batch x [B, T, F] (numpy fp32); returns a masked copy."""
    x = np.array(x, dtype=np.float32, copy=True)
    for spec in x:  # views of the padded batch: tau = T_max
        if random.random() < rate:
            tau, v = spec.shape
            for _ in range(frequency_masking_num):  # specaugment.py:97-101
                f = int(np.random.uniform(low=0.0, high=frequency_masking_pars))
                f0 = random.randint(0, v - f)
                spec[:, f0:f0 + f] = 0
            for _ in range(time_masking_num):  # specaugment.py:104-108
                t = int(np.random.uniform(low=1.0, high=min(time_masking_pars, tau)))
                t0 = random.randint(0, tau - t)
                spec[t0:t0 + t, :] = 0
    return x


def linspace_round(start, end, steps):
    """round(torch.linspace(start, end, steps)) as int64 (fp32 arithmetic, half-to-even)."""
    if steps <= 0:
        return np.zeros(0, dtype=np.int64)
    s, e = np.float32(start), np.float32(end)
    if steps == 1:
        return np.array([int(start)], dtype=np.int64)
    step = np.float32(np.float32(e - s) / np.float32(steps - 1))
    i = np.arange(steps)
    up = (s + (step * i.astype(np.float32)).astype(np.float32)).astype(np.float32)
    down = (e - (step * (steps - 1 - i).astype(np.float32)).astype(np.float32)).astype(np.float32)
    return np.rint(np.where(i < steps // 2, up, down)).astype(np.int64)


def time_stretch_ids(time_len, w, low=0.8, high=1.25):
    """time_stretch.py:40-54: source-frame indices of the stretched utterance."""
    if time_len < 10 and low < 1.0:
        low = 1.0
    ids = []
    for i in range(int(round(time_len / w))):
        s = random.uniform(low, high) * min(w, time_len - w * i)
        e = min(time_len, w * (i + 1))
        ids.append(linspace_round(w * i, e - 1, int(s)))
    return np.concatenate(ids) if ids else np.zeros(0, dtype=np.int64)


def time_stretch_batch(x, lengths, rate, w, low, high):
    """time_stretch.py:18-38: returns (frames [B, max new length, F] fp32, new lengths, ids per utterance)."""
    x = np.asarray(x, dtype=np.float32)
    ids_all = []
    for b, n in enumerate(lengths):
        if random.random() < rate:
            ids_all.append(time_stretch_ids(int(n), w, low, high))
        else:
            ids_all.append(np.arange(int(n), dtype=np.int64))
    new_lengths = [len(i) for i in ids_all]
    frames = np.zeros((len(lengths), max(new_lengths), x.shape[2]), dtype=np.float32)
    for b, ids in enumerate(ids_all):
        frames[b, :len(ids)] = x[b, ids]
    return frames, new_lengths, ids_all
