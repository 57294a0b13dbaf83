"""Batch synthesis harness for the diffusion decoder: host batches in, host mels out, with the host<->device copies
of neighbouring batches overlapped with the reverse diffusion of the current one (three CUDA streams, double
buffering).  This is the end-to-end path ``bench.py`` times as ``e2e``; the reference's equivalent is the batch loop
of ``synthesize.py:106-140`` (``to_device`` -> model -> ``.cpu()``), which is strictly serial.

One process drives one GPU; multi-GPU synthesis runs one ``BatchSynthesizer`` per rank on its utterance shard
(``mixgan_tts_b200.shard``) with no collective.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional, Tuple

import torch

from .diffusion import GaussianDiffusion


class BatchSynthesizer:
    """``for mel in BatchSynthesizer(gd).run(batches): ...`` where every batch is a tuple
    ``(cond [B,T,H] fp32, mel_mask [B,T] bool (True = padding), spk [B,H] | None, coarse_mel [B,T,M] | None)`` of
    PINNED host tensors and every yielded ``mel`` is a pinned host ``[B,T,M]`` tensor (valid until the next two
    batches have been yielded: the output buffers are recycled)."""

    def __init__(self, diffusion: GaussianDiffusion, device: Optional[torch.device] = None, aux_decoder=None, vocoder=None):
        """``aux_decoder`` (``mixgan_tts_b200.AuxDecoder``): batches may then carry ``coarse_mel = None`` in `shallow` mode
        and the coarse mel is computed on the GPU from ``cond`` (which IS the decoder's input: model/mixgantts.py:138-139).
        ``vocoder`` (``mixgan_tts_b200.Generator``): the yielded tensors are waveforms ``[B, T * hop]`` instead of mels."""
        self.gd = diffusion
        self.aux, self.voc = aux_decoder, vocoder
        self.device = device or next(diffusion.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("BatchSynthesizer needs the model on a CUDA device (no CPU fallback)")
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self._out = {}

    def _upload(self, batch):
        with torch.cuda.stream(self.s_in):
            dev = tuple(None if t is None else t.to(self.device, non_blocking=True) for t in batch)
            ev = torch.cuda.Event()
            ev.record(self.s_in)
        return dev, ev

    def _host_out(self, slot: int, shape) -> torch.Tensor:
        buf = self._out.get(slot)
        if buf is None or tuple(buf.shape) != tuple(shape):
            buf = torch.empty(shape, dtype=torch.float32).pin_memory()
            self._out[slot] = buf
        return buf

    @torch.no_grad()
    def run(self, batches: Iterable[Tuple]) -> Iterator[torch.Tensor]:
        compute = torch.cuda.current_stream(self.device)
        it = iter(batches)
        nxt = next(it, None)
        if nxt is None:
            return
        pending = self._upload(tuple(nxt) + (None,) * (4 - len(nxt)))
        done = []                                    # (host mel, event) of batches whose D2H copy is in flight
        i = 0
        while pending is not None:
            (cond, mask, spk, coarse), ev_in = pending
            nxt = next(it, None)
            pending = self._upload(tuple(nxt) + (None,) * (4 - len(nxt))) if nxt is not None else None
            compute.wait_event(ev_in)                # inputs of THIS batch have landed; the next upload runs meanwhile
            if coarse is None and self.aux is not None and self.gd.model == "shallow":
                coarse = self.aux(cond, mask)        # model/mixgantts.py:139-143, stays on the device
            mel = self.gd(None, cond, spk, mask, coarse_mel=coarse)[0]
            if self.voc is not None:
                mel = self.voc.forward_frames(mel)   # utils/model.py:103-110
            for t in (cond, mask, spk, coarse):      # the upload stream allocated them; the compute stream used them
                if t is not None:
                    t.record_stream(compute)
            ev_c = torch.cuda.Event()
            ev_c.record(compute)
            host = self._host_out(i % 3, mel.shape)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_c)
                host.copy_(mel, non_blocking=True)
                mel.record_stream(self.s_out)
                ev_o = torch.cuda.Event()
                ev_o.record(self.s_out)
            done.append((host, ev_o))
            if len(done) > 1:                        # hand out batch i-1 while batch i computes
                h, e = done.pop(0)
                e.synchronize()
                yield h
            i += 1
        for h, e in done:
            e.synchronize()
            yield h
