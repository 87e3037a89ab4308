"""Ragged-utterance staging: the step *before* the scoring path.

The reference fits every utterance to ``test_duration_sec * sample_rate`` samples on the CPU inside the
dataset (``adjustDuration`` / ``adjustDuration_random_start``, reference data/test_set.py:201-248: tile-repeat
short utterances, crop long ones from the start or from ``random.randint(0, len - duration)``), collates
(B, duration) fp32 batches in DataLoader workers and copies them from pageable memory (main.py:200-209).

Here the host only packs what is needed -- at most ``duration`` samples per utterance, the whole utterance
when it is shorter -- into pinned memory; one async H2D copy per batch runs on a side stream, and the
``fit_duration`` kernel of librtdf.so tiles / crops (and optionally pre-emphasises, data/preprocess.py:22-27)
in HBM.  Two slots are double-buffered so packing batch k+1 overlaps the forward of batch k.

PyTorch supplies pinned/device memory, streams and events; there is no CPU implementation of the fit.
"""
import random as _random

import numpy as np
import torch

from .rtdf_runtime import native


def crop_starts(lengths, duration, random_start=False, rng=_random):
    """Start offsets per utterance.  random_start follows the reference draw for draw
    (data/test_set.py:229-246): one ``randint(0, len(x) - duration)`` per utterance, where a tiled utterance has
    exactly ``duration`` samples (so the draw is ``randint(0, 0)`` and still consumes the generator)."""
    starts = []
    for n in lengths:
        if n < 1:
            raise ValueError("empty utterance")
        fitted = duration if n < duration else n
        starts.append(rng.randint(0, fitted - duration) if random_start else 0)
    return starts


class UtteranceStager:
    """Double-buffered pinned staging of ragged utterance lists into fixed-length device batches."""

    def __init__(self, duration, batch_size, device, preemph=None, random_start=False, rng=None, slots=2):
        self.lib = native.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("UtteranceStager stages into CUDA memory only (no CPU path)")
        self.duration, self.batch_size = int(duration), int(batch_size)
        self.preemph = None if preemph is None else float(preemph)
        self.random_start = bool(random_start)
        self.rng = rng or _random
        cap = self.batch_size * self.duration
        self.slots = [{
            "host": torch.empty(cap, dtype=torch.float32).pin_memory(),
            "host_off": torch.empty(self.batch_size + 1, dtype=torch.int64).pin_memory(),
            "host_start": torch.empty(self.batch_size, dtype=torch.int32).pin_memory(),
            "dev": torch.empty(cap, dtype=torch.float32, device=self.device),
            "dev_off": torch.empty(self.batch_size + 1, dtype=torch.int64, device=self.device),
            "dev_start": torch.empty(self.batch_size, dtype=torch.int32, device=self.device),
            "out": torch.empty(self.batch_size, self.duration, dtype=torch.float32, device=self.device),
            "copied": torch.cuda.Event(), "free": torch.cuda.Event(), "outstanding": False,
        } for _ in range(slots)]
        self.copy_stream = torch.cuda.Stream(self.device)
        self.cur = 0
        self.h2d_bytes = 0

    def stage(self, utterances):
        """utterances: list of 1-D (or (1,n)) float arrays / CPU tensors.  Packs, starts the H2D copy on the side
        stream and returns a ticket for ``take``."""
        B = len(utterances)
        if not 1 <= B <= self.batch_size:
            raise ValueError(f"batch of {B} utterances (1..{self.batch_size} supported)")
        slot = self.slots[self.cur]
        self.cur = (self.cur + 1) % len(self.slots)
        if slot["outstanding"]:
            raise RuntimeError("UtteranceStager: slot re-staged while its previous ticket is outstanding -- call "
                               "release(ticket) after the forward that consumed take(ticket) (or raise `slots`)")
        slot["free"].synchronize()              # the forward that consumed this slot's `out` has finished
        slot["outstanding"] = True
        flat = [np.asarray(u, dtype=np.float32).reshape(-1) if not torch.is_tensor(u)
                else u.detach().to(torch.float32).reshape(-1).numpy() for u in utterances]
        starts = crop_starts([len(u) for u in flat], self.duration, self.random_start, self.rng)
        host = slot["host"].numpy()
        off = 0
        offs = [0]
        for u, st in zip(flat, starts):
            # long utterances: only the cropped window crosses PCIe; short ones: the utterance once, tiled in HBM
            seg = u[st:st + self.duration] if len(u) >= self.duration else u
            host[off:off + len(seg)] = seg
            off += len(seg)
            offs.append(off)
        slot["host_off"][:B + 1] = torch.tensor(offs, dtype=torch.int64)
        slot["host_start"][:B] = 0              # the crop already happened on the host side of the copy
        with torch.cuda.stream(self.copy_stream):
            slot["dev"][:off].copy_(slot["host"][:off], non_blocking=True)
            slot["dev_off"][:B + 1].copy_(slot["host_off"][:B + 1], non_blocking=True)
            slot["dev_start"][:B].copy_(slot["host_start"][:B], non_blocking=True)
            slot["copied"].record(self.copy_stream)
        self.h2d_bytes += off * 4 + (B + 1) * 8 + B * 4
        return slot, B

    def take(self, ticket):
        """(B, duration) fp32 device batch for a staged ticket (fit kernel on the current stream)."""
        slot, B = ticket
        stream = torch.cuda.current_stream(self.device)
        stream.wait_event(slot["copied"])
        with torch.cuda.device(self.device):
            native.check(self.lib.rtdf_fit_duration(
                native.ptr(slot["dev"]), native.ptr(slot["dev_off"]), native.ptr(slot["dev_start"]), B, self.duration,
                int(self.preemph is not None), float(self.preemph or 0.0), native.ptr(slot["out"]),
                stream.cuda_stream), "rtdf_fit_duration")
        return slot["out"][:B]

    def release(self, ticket):
        """Call after enqueueing the work that reads ``take(ticket)`` (same stream): records the event the next
        ``stage`` into this slot waits for, so the slot's pinned / device buffers are not overwritten while the copy,
        the fit kernel or the forward may still be reading them.  A slot whose ticket was not released cannot be
        staged again (``stage`` raises)."""
        ticket[0]["free"].record(torch.cuda.current_stream(self.device))
        ticket[0]["outstanding"] = False


def fit_duration(packed, offsets, duration, starts=None, preemph=None):
    """Device-side fit of already resident ragged data: packed fp32 (sum len), offsets int64 (B+1) ->
    (B, duration) fp32.  Thin wrapper over ``rtdf_fit_duration`` (used by the parity tests)."""
    if not packed.is_cuda:
        raise RuntimeError("fit_duration: CUDA tensors only (no CPU path)")
    B = offsets.numel() - 1
    out = torch.empty(B, int(duration), dtype=torch.float32, device=packed.device)
    with torch.cuda.device(packed.device):
        native.check(native.load().rtdf_fit_duration(
            native.ptr(packed), native.ptr(offsets), native.ptr(starts), B, int(duration),
            int(preemph is not None), float(preemph or 0.0), native.ptr(out),
            torch.cuda.current_stream(packed.device).cuda_stream), "rtdf_fit_duration")
    return out
