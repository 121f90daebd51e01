"""Host-side partitioning of utterances across GPUs: length-bucketed batches, one
process per GPU, no collective on the forward path.

The reference builds batches with ``batch_by_size`` over UNSORTED indices
(``fairseq/data/data_utils_fast.pyx:27-68``; ``data/fbank_dataset.py:78-81``) and deals them
round-robin to ranks (``fairseq/data/iterators.py:383-413``), so a step waits for whichever rank drew
the longest batch.  Here utterances are sorted by length, cut into batches of at most ``max_frames``
PADDED frames (``len(batch) * longest``, the quantity the kernels actually process) and every group of
``world`` consecutive batches -- which have near-identical cost -- forms one step, one batch per rank.
"""
from typing import List, Sequence


def padded_length(n: int, pad_multiple: int = 1) -> int:
    """Frames the batch tensor is padded to: the longest utterance rounded up to ``pad_multiple``."""
    return (n + pad_multiple - 1) // pad_multiple * pad_multiple if pad_multiple > 1 else n


def bucket_by_length(lengths: Sequence[int], max_frames: int, max_sentences: int = 0,
                     pad_multiple: int = 1) -> List[List[int]]:
    """Batches of utterance indices, longest first inside a batch (collater order,
    ``data/collaters.py:89-92``); ``len(batch) * padded(max(len)) <= max_frames``.

    ``pad_multiple`` > 1 quantises the batch's time extent (T is rounded up to a multiple of it), so the
    set of distinct batch SHAPES of an epoch is small: a ragged stream then re-uses the encoder's per-shape
    CUDA graphs instead of re-capturing on every step (every batch of a given padded T is filled to the same
    ``max_frames // T`` utterances, except the last one)."""
    if max_frames <= 0:
        raise ValueError("max_frames must be positive")
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    batches, cur = [], []
    for i in order:
        n = padded_length(lengths[i], pad_multiple)
        if n > max_frames:
            raise ValueError("utterance %d has %d frames > max_frames=%d" % (i, n, max_frames))
        longest = padded_length(lengths[cur[0]], pad_multiple) if cur else n
        if cur and ((len(cur) + 1) * longest > max_frames or (max_sentences and len(cur) >= max_sentences)):
            batches.append(cur)
            cur = []
        cur.append(i)
    if cur:
        batches.append(cur)
    return batches


def shard_steps(batches: List[List[int]], world: int) -> List[List[List[int]]]:
    """steps[s][rank] -> batch.  Consecutive (similar-cost) batches share a step; the tail step is
    padded with empty batches so that every rank runs the same number of steps."""
    steps = []
    for s in range(0, len(batches), world):
        group = batches[s:s + world]
        steps.append(group + [[] for _ in range(world - len(group))])
    return steps


def batches_for_rank(lengths: Sequence[int], max_frames: int, rank: int, world: int,
                     max_sentences: int = 0) -> List[List[int]]:
    steps = shard_steps(bucket_by_length(lengths, max_frames, max_sentences), world)
    return [st[rank] for st in steps]


def step_imbalance(lengths: Sequence[int], steps: List[List[List[int]]], pad_multiple: int = 1) -> float:
    """max over steps of (largest padded-frame count) / (mean padded-frame count) over busy ranks."""
    worst = 1.0
    for st in steps:
        cost = [len(b) * padded_length(max(lengths[i] for i in b), pad_multiple) for b in st if b]
        if len(cost) > 1:
            worst = max(worst, max(cost) / (sum(cost) / len(cost)))
    return worst
