"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference collater for the "next" row N2.

Follows ``Seq2SeqCollater`` (examples/speech_recognition/data/collaters.py:43-131) and
``collate_tokens`` (fairseq/data/data_utils.py:33-48), plus ``apply_mv_norm`` per utterance
(data/data_utils.py:9-24 via ``encoder_oracle.cmvn``) for the normalising variant.

Pinned by the reference's own known-answer test, tests/speech_recognition/test_collaters.py:23-50
(``tests/test_oracle.py::test_collate_oracle_reference_kat``), and by the live reference class when
/root/reference is mounted.  Only tests/ may import this module.
"""
import numpy as np
import torch

from . import encoder_oracle as O


def collate_frames(frames):
    """collaters.py:43-56."""
    len_max = max(f.size(0) for f in frames)
    res = frames[0].new_zeros(len(frames), len_max, frames[0].size(1))
    for i, v in enumerate(frames):
        res[i, : v.size(0)] = v
    return res


def collate_tokens(values, pad_idx, eos_idx, move_eos_to_beginning):
    """fairseq/data/data_utils.py:33-48 (left_pad=False)."""
    size = max(v.size(0) for v in values)
    res = values[0].new_full((len(values), size), pad_idx)
    for i, v in enumerate(values):
        if move_eos_to_beginning:
            res[i, 0] = eos_idx
            res[i, 1:len(v)] = v[:-1]
        else:
            res[i, :len(v)] = v
    return res


def collate(samples, pad_index=1, eos_index=2, move_eos_to_beginning=True, normalize=False):
    """collaters.py:58-131 with feature_index=0, label_index=1."""
    src = [torch.from_numpy(s["data"][0]) if isinstance(s["data"][0], np.ndarray) else s["data"][0]
           for s in samples]
    if normalize:
        src = [O.cmvn(f.float()) for f in src]
    tgt = [torch.as_tensor(s["data"][1]).long() for s in samples]
    ids = torch.LongTensor([s["id"] for s in samples])
    frames = collate_frames(src)
    lengths = torch.LongTensor([f.size(0) for f in src])
    lengths, order = lengths.sort(descending=True)
    return {
        "id": ids.index_select(0, order),
        "ntokens": sum(len(t) for t in tgt),
        "net_input": {
            "src_tokens": frames.index_select(0, order),
            "src_lengths": lengths,
            "prev_output_tokens": collate_tokens(tgt, pad_index, eos_index, move_eos_to_beginning)
            .index_select(0, order),
        },
        "target": collate_tokens(tgt, pad_index, eos_index, False).index_select(0, order),
        "target_lengths": torch.LongTensor([len(t) for t in tgt]).index_select(0, order),
        "nsentences": len(samples),
    }
