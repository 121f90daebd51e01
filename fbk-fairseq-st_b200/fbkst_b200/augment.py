"""Device-side SpecAugment and TimeStretch.

Mirrors of ``examples/speech_recognition/modules/specaugment.py:42-67`` (``SpecAugment``) and
``modules/time_stretch.py:7-38`` (``TimeStretch``): same constructor arguments, same
``forward(batch)`` contract (``batch['net_input']['src_tokens']`` B x T x F and, for TimeStretch,
``['src_lengths']``), called by the task right before the criterion
(``tasks/speech_recognition.py:254-258``).

What stays on the host are the random draws: they are a few numbers per utterance and they consume
Python's ``random`` and ``numpy.random`` streams with the same calls in the same order as the
reference, so under the same seeds the masks and the stretch factors are the reference's.  What
moves to the device is everything that touches frames: the reference masks sample by sample and
re-stacks the batch (SpecAugment), or deep-copies the batch, fancy-indexes every utterance on the
host, re-pads and ships the whole batch to the GPU a second time (TimeStretch).  Here one small
int32 table goes up and one kernel masks in place / resamples and re-pads.

There is no CPU path: without the CUDA library / a B200 ``forward`` raises.
"""
import random

import numpy as np
import torch

from . import ops


class SpecAugment(torch.nn.Module):
    """specaugment.py:42-67; masks ``src_tokens`` in place (the reference's slice assignments also
    write through to the batch tensor before it is re-stacked)."""

    def __init__(self, frequency_masking_pars, time_masking_pars, frequency_masking_num, time_masking_num,
                 rate=1.0):
        super().__init__()
        self.frequency_masking_pars = frequency_masking_pars
        self.time_masking_pars = time_masking_pars
        self.frequency_masking_num = frequency_masking_num
        self.time_masking_num = time_masking_num
        self.rate = rate

    def draw_bands(self, B, tau, v):
        """[B, m_F + m_T, 2] int32 (start, width), drawn like specaugment.py:56-58 then :97-109 for
        each spectrogram of the PADDED batch (tau = T_max, v = F)."""
        nf, nt = self.frequency_masking_num, self.time_masking_num
        bands = np.zeros((B, nf + nt, 2), dtype=np.int32)
        for b in range(B):
            if random.random() < self.rate:
                for i in range(nf):
                    f = int(np.random.uniform(low=0.0, high=self.frequency_masking_pars))
                    bands[b, i] = (random.randint(0, v - f), f)
                for i in range(nt):
                    t = int(np.random.uniform(low=1.0, high=min(self.time_masking_pars, tau)))
                    bands[b, nf + i] = (random.randint(0, tau - t), t)
        return bands

    def forward(self, batch):
        x = batch["net_input"]["src_tokens"]
        if not x.is_cuda:
            raise RuntimeError("fbkst_b200.SpecAugment: src_tokens must be on the GPU (no CPU path)")
        B, tau, v = x.shape
        bands = torch.from_numpy(self.draw_bands(B, tau, v)).to(x.device, non_blocking=True)
        if not x.is_contiguous():
            x = x.contiguous()
        ops.specaugment_(x, bands, self.frequency_masking_num, self.time_masking_num)
        batch["net_input"]["src_tokens"] = x
        return batch


def stretch_windows(time_len, w, low, high):
    """The windows of time_stretch_seq (time_stretch.py:40-54) for one utterance as int arrays
    (first, last, count): window i resamples frames w*i .. min(time_len, w*(i+1)) - 1 to
    int(uniform(low, high) * min(w, time_len - w*i)) frames.  Consumes ``random.uniform`` once per
    window, like the reference."""
    if time_len < 10 and low < 1.0:
        low = 1.0
    n = int(round(time_len / w))
    draws = np.array([random.uniform(low, high) for _ in range(n)], dtype=np.float64)
    i = np.arange(n, dtype=np.int64)
    first = w * i
    size = np.minimum(w, time_len - first)
    count = (draws * size).astype(np.int64)  # int(s): truncation
    last = np.minimum(time_len, first + w) - 1
    if n and (count.min() < 0 or (last < first).any()):
        raise ValueError("fbkst_b200.TimeStretch: negative window size (low must be positive)")
    return first, last, count


class TimeStretch(torch.nn.Module):
    """time_stretch.py:7-38."""

    def __init__(self, rate, w, low, high):
        super().__init__()
        if w < 1:
            raise ValueError("w must be greater than 1")
        self.w = w
        self.low = low
        self.high = high
        self.rate = rate
        self.last_ids = None  # [B, T_out] int32 source frame of every output frame (-1 = padding)

    def draw_windows(self, lengths):
        """Per-utterance windows + new lengths.  Returns (windows [n,4] int64 with out_off still
        relative to the utterance, utterance index per window, new_lengths list)."""
        firsts, lasts, counts, offs, owner, new_lengths = [], [], [], [], [], []
        for b, length in enumerate(lengths):
            if random.random() < self.rate:
                first, last, count = stretch_windows(length, self.w, self.low, self.high)
            else:  # elem[:length] unchanged = one identity window
                first = np.zeros(1, dtype=np.int64)
                last = np.full(1, length - 1, dtype=np.int64)
                count = np.full(1, length, dtype=np.int64)
            off = np.cumsum(count) - count
            firsts.append(first); lasts.append(last); counts.append(count); offs.append(off)
            owner.append(np.full(len(count), b, dtype=np.int64))
            new_lengths.append(int(count.sum()))
        cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, dtype=np.int64)  # noqa: E731
        return np.stack([cat(firsts), cat(lasts), cat(counts), cat(offs)], axis=1), cat(owner), new_lengths

    def forward(self, batch):
        x = batch["net_input"]["src_tokens"]
        if not x.is_cuda:
            raise RuntimeError("fbkst_b200.TimeStretch: src_tokens must be on the GPU (no CPU path)")
        lengths = [int(n) for n in batch["net_input"]["src_lengths"].tolist()]
        windows, owner, new_lengths = self.draw_windows(lengths)
        T_out = max(max(new_lengths), 1)
        windows[:, 3] += owner * T_out  # flat offset into [B, T_out]
        keep = windows[:, 2] > 0
        wdev = torch.from_numpy(np.ascontiguousarray(windows[keep].astype(np.int32))).to(x.device,
                                                                                          non_blocking=True)
        if wdev.numel() == 0:
            wdev = torch.zeros(1, 4, dtype=torch.int32, device=x.device)[:0]
        out, ids = ops.time_stretch(x.contiguous(), wdev, T_out)
        if max(new_lengths) == 0:
            out = out[:, :0]
            ids = ids[:, :0]
        self.last_ids = ids
        new_batch = dict(batch)
        new_batch["net_input"] = dict(batch["net_input"])
        new_batch["net_input"]["src_tokens"] = out
        new_batch["net_input"]["src_lengths"] = torch.tensor(new_lengths, dtype=torch.long).to(x.device)
        return new_batch
