"""Host-buffer serving loop for the encoder: the public end-to-end entry point.

``EncoderPipeline.run`` takes batches that live in (pinned) HOST memory and returns encoder outputs in
HOST memory.  Three CUDA streams overlap the host->device copy of batch i+1, the kernels of batch i
(fbank CMVN + encoder forward) and the device->host copy of batch i-1; all arithmetic still happens
in the sm_100a kernels, PyTorch only moves bytes and orders streams.

The per-batch work is exactly ``apply_mv_norm`` (data/fbank_dataset.py:44-45) followed by
``ConvolutionalTransformerEncoder.forward`` (models/conv_transformer.py:195-276).
"""
import collections

import torch

from . import ops


class EncoderPipeline:
    def __init__(self, encoder, normalize=True, device=None):
        self.enc = encoder
        self.normalize = normalize
        self.device = device or next(encoder.parameters()).device
        self.s_in = torch.cuda.Stream(self.device)
        self.s_compute = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self._dev_in = {}    # (slot, shape) -> device staging buffer
        self._host_out = {}  # (slot, shape) -> pinned host buffer
        self._busy = [None, None]  # event: compute finished reading input slot

    def _in_buffer(self, slot, shape):
        key = (slot, tuple(shape))
        if key not in self._dev_in:
            self._dev_in[key] = torch.empty(shape, dtype=torch.float32, device=self.device)
        return self._dev_in[key]

    def _out_buffer(self, slot, shape, dtype):
        key = (slot, tuple(shape), dtype)
        if key not in self._host_out:
            self._host_out[key] = torch.empty(shape, dtype=dtype).pin_memory()
        return self._host_out[key]

    def run(self, batches):
        """``batches``: iterable of ``(src_tokens [B,T,F] fp32 host tensor, src_lengths [B] int64 host
        tensor)``.  Yields ``(encoder_out [T'',B,D] fp32 host tensor, out_lengths [B] host tensor)`` in
        order.  A yielded tensor is a view of a reusable pinned buffer: consume it before asking for
        the batch after next."""
        pending = collections.deque()
        staged = None
        it = iter(batches)

        def stage(i, batch):
            x_host, lengths = batch
            buf = self._in_buffer(i & 1, x_host.shape)
            with torch.cuda.stream(self.s_in):
                if self._busy[i & 1] is not None:
                    self.s_in.wait_event(self._busy[i & 1])
                buf.copy_(x_host, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.s_in)
            return buf, lengths, ev

        i = 0
        first = next(it, None)
        if first is None:
            return
        staged = stage(0, first)
        while staged is not None:
            x_dev, lengths, ev_in = staged
            nxt = next(it, None)
            staged = stage(i + 1, nxt) if nxt is not None else None  # H2D(i+1) overlaps compute(i)
            with torch.cuda.stream(self.s_compute):
                self.s_compute.wait_event(ev_in)
                len32 = lengths.to(torch.int32).to(self.device, non_blocking=True)
                x = ops.cmvn(x_dev, len32) if self.normalize else x_dev
                out = self.enc(x, lengths)
                ev_done = torch.cuda.Event()
                ev_done.record(self.s_compute)
            self._busy[i & 1] = ev_done
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_done)
                eo = out.encoder_out
                host = self._out_buffer(i & 1, eo.shape, eo.dtype)
                host.copy_(eo, non_blocking=True)
                hl = self._out_buffer(i & 1, out.src_lengths.shape, out.src_lengths.dtype)
                hl.copy_(out.src_lengths, non_blocking=True)
                ev_out = torch.cuda.Event()
                ev_out.record(self.s_out)
            pending.append((ev_out, host, hl, out))
            if len(pending) > 1:  # D2H(i-1) has been overlapping compute(i)
                e, h, l, _keep = pending.popleft()
                e.synchronize()
                yield h, l
            i += 1
        while pending:
            e, h, l, _keep = pending.popleft()
            e.synchronize()
            yield h, l
