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


def bucket_by_length(lengths: Sequence[int], max_frames: int, max_sentences: int = 0) -> List[List[int]]:
    """Batches of utterance indices, longest first inside a batch (collater order,
    ``data/collaters.py:89-92``); ``len(batch) * max(len) <= max_frames``."""
    if max_frames <= 0:
        raise ValueError("max_frames must be positive")
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    batches, cur = [], []
    for i in order:
        n = lengths[i]
        if n > max_frames:
            raise ValueError("utterance %d has %d frames > max_frames=%d" % (i, n, max_frames))
        longest = lengths[cur[0]] if cur else n
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


def step_imbalance(lengths: Sequence[int], steps: List[List[List[int]]]) -> float:
    """max over steps of (largest padded-frame count) / (mean padded-frame count) over busy ranks."""
    worst = 1.0
    for st in steps:
        cost = [len(b) * max(lengths[i] for i in b) for b in st if b]
        if len(cost) > 1:
            worst = max(worst, max(cost) / (sum(cost) / len(cost)))
    return worst
