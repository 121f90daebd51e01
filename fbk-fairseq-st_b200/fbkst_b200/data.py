"""Device collater: the batch-assembly step right before the encoder.

Mirror of ``Seq2SeqCollater`` (examples/speech_recognition/data/collaters.py:21-131): same
constructor arguments, same ``collate(samples)`` result (``id``, ``ntokens``, ``nsentences``,
``net_input = {src_tokens, src_lengths, prev_output_tokens}``, ``target``, ``target_lengths``),
samples sorted by descending frame count (:89-92).  What changes is where the frames are padded:

* the reference pads on the host (``_collate_frames`` :43-56) and ships B x T_max x F floats;
* here the utterances are packed back to back into ONE pinned buffer (no padding bytes), copied
  with one H2D, and ``fbkst_collate_cmvn_f32`` writes the zero-padded ``src_tokens`` on the device,
  optionally applying the per-utterance fbank CMVN (data/data_utils.py:9-24, which the reference
  dataset does on the host per sample, data/fbank_dataset.py:44-45) in the same pass.

Targets are tiny integer tensors: they are collated on the host exactly like
``fairseq.data.data_utils.collate_tokens`` (fairseq/data/data_utils.py:33-48).  There is no CPU path
for the frames: without the CUDA library / a B200 ``collate`` raises.
"""
import numpy as np
import torch

from . import ops


def collate_tokens(values, pad_idx, eos_idx=None, left_pad=False, move_eos_to_beginning=False):
    """fairseq/data/data_utils.py:33-48."""
    size = max(v.size(0) for v in values)
    res = values[0].new_full((len(values), size), pad_idx)
    for i, v in enumerate(values):
        dst = res[i][size - len(v):] if left_pad else res[i][:len(v)]
        if move_eos_to_beginning:
            dst[0] = eos_idx
            dst[1:] = v[:-1]
        else:
            dst.copy_(v)
    return res


class DeviceCollater:
    def __init__(self, feature_index=0, label_index=1, pad_index=1, eos_index=2,
                 move_eos_to_beginning=True, normalize=False, device="cuda:0"):
        self.feature_index = feature_index
        self.label_index = label_index
        self.pad_index = pad_index
        self.eos_index = eos_index
        self.move_eos_to_beginning = move_eos_to_beginning
        self.normalize = normalize
        self.device = torch.device(device)
        self._pinned = None

    def _pack(self, frames):
        """Utterances back to back in one pinned fp32 buffer.  Returns (view, release): the caller
        enqueues the H2D copy of ``view`` and then calls ``release()``, which records a CUDA event
        behind the copy; a pinned buffer is reused only once that event has completed, so a host
        that runs ahead of the GPU never overwrites a batch whose copy has not executed yet."""
        total = sum(f.shape[0] for f in frames)
        fdim = frames[0].shape[1]
        need = max(total * fdim, 1)
        if self._pinned is None:
            self._pinned = []  # [buffer, event or None]
        slot = None
        for entry in self._pinned:
            if entry[0].numel() >= need and (entry[1] is None or entry[1].query()):
                slot = entry
                break
        if slot is None:
            slot = [torch.empty(need, dtype=torch.float32).pin_memory(), None]
            self._pinned.append(slot)
            if len(self._pinned) > 8:  # bounded: drop a finished, smaller buffer
                for i, entry in enumerate(self._pinned[:-1]):
                    if entry[1] is None or entry[1].query():
                        self._pinned.pop(i)
                        break
        buf = slot[0][: total * fdim].view(total, fdim)
        o = 0
        for f in frames:
            buf[o:o + f.shape[0]].copy_(f)
            o += f.shape[0]

        def release():
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            slot[1] = ev
        return buf, release

    def collate_frames(self, frames, order=None):
        """``frames``: list of [T_i, F] tensors/arrays.  Returns (src_tokens [B,T_max,F] on the device
        in slot order ``order`` (default: descending length), src_lengths [B] int64 host, order)."""
        frames = [torch.from_numpy(f) if isinstance(f, (np.ndarray, np.generic)) else f for f in frames]
        frames = [f.float() for f in frames]
        lens = torch.tensor([f.shape[0] for f in frames], dtype=torch.long)
        if order is None:
            lens_sorted, order = lens.sort(descending=True)  # collaters.py:89-90
        else:
            lens_sorted = lens.index_select(0, order)
        starts_all = torch.cumsum(lens, 0) - lens
        staged, release = self._pack(frames)
        packed = staged.to(self.device, non_blocking=True)
        release()  # event behind the H2D copy guards the pinned buffer
        starts = starts_all.index_select(0, order).to(self.device, non_blocking=True)
        len32 = lens_sorted.to(torch.int32).to(self.device, non_blocking=True)
        src = ops.collate_cmvn(packed, starts, len32, int(lens_sorted.max()), normalize=self.normalize)
        return src, lens_sorted, order

    def collate(self, samples):
        if len(samples) == 0:
            return {}
        parsed = []
        for s in samples:  # collaters.py:66-82
            source = s["data"][self.feature_index]
            if source is None:
                continue
            target = s["data"][self.label_index]
            if isinstance(target, (np.ndarray, np.generic)):
                target = torch.from_numpy(target).long()
            elif isinstance(target, list):
                target = torch.LongTensor(target)
            parsed.append({"id": s["id"], "source": source, "target": target})
        samples = parsed
        ids = torch.LongTensor([s["id"] for s in samples])
        frames, frames_lengths, sort_order = self.collate_frames([s["source"] for s in samples])
        ids = ids.index_select(0, sort_order)
        target = target_lengths = prev_output_tokens = None
        if samples[0].get("target", None) is not None:  # collaters.py:97-119
            ntokens = sum(len(s["target"]) for s in samples)
            target = collate_tokens([s["target"] for s in samples], self.pad_index, self.eos_index,
                                    left_pad=False, move_eos_to_beginning=False)
            target = target.index_select(0, sort_order)
            target_lengths = torch.LongTensor([s["target"].size(0) for s in samples]).index_select(0, sort_order)
            prev_output_tokens = collate_tokens([s["target"] for s in samples], self.pad_index,
                                                self.eos_index, left_pad=False,
                                                move_eos_to_beginning=self.move_eos_to_beginning)
            prev_output_tokens = prev_output_tokens.index_select(0, sort_order)
        else:
            ntokens = sum(len(s["source"]) for s in samples)
        batch = {"id": ids, "ntokens": ntokens,
                 "net_input": {"src_tokens": frames, "src_lengths": frames_lengths},
                 "target": target, "target_lengths": target_lengths, "nsentences": len(samples)}
        if prev_output_tokens is not None:
            batch["net_input"]["prev_output_tokens"] = prev_output_tokens
        return batch
