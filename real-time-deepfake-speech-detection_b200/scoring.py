"""Data-parallel scoring driver: utterances sharded across the GPUs of one box, one collective.

Mirrors what the reference's scoring callers do per batch -- ``produce_evaluation_file``
(reference main.py:199-221) and ``Trainer._test`` (trainer.py:85-132): ``model(batch_x)`` under
``no_grad`` and ``score = out[:, 1]`` -- but

  * rank r of W scores the contiguous index range [r*S, min((r+1)*S, n)), S = ceil(n/W)
    (SURVEY.md section 8e); every rank holds a full weight replica, there is no data-path collective;
  * scores stay on the device and are exchanged with ONE all-gather of fp32[S] per rank
    (NCCL over NVLink on GPUs; gloo in the CPU tests of the host logic), tail padded with NaN;
  * H2D copies come from pinned, double-buffered host memory on a side stream.

PyTorch supplies device memory, streams and torch.distributed; the forward itself is librtdf.so.
"""
import math

import torch
import torch.distributed as dist

PAD = float("nan")


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of rank `rank`; S = ceil(n/W) items per rank except the tail."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    per = math.ceil(n_items / world) if n_items > 0 else 0
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items), per


def batch_ranges(lo, hi, batch_size):
    """[(b_lo, b_hi)] covering [lo, hi) in order; the last batch is ragged (drop_last=False, main.py:200)."""
    if batch_size < 1:
        raise ValueError("batch_size must be >= 1")
    return [(s, min(s + batch_size, hi)) for s in range(lo, hi, batch_size)]


def gather_scores(local_scores, n_items, per_rank, group=None):
    """All-gather fixed-size, NaN-padded shards and return the first n_items scores in global order.
    local_scores: 1-D fp32 tensor with this rank's (hi-lo) scores, on the device the backend needs
    (CUDA for nccl, CPU for gloo)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    send = torch.full((per_rank,), PAD, dtype=torch.float32, device=local_scores.device)
    send[: local_scores.numel()] = local_scores
    if world == 1:
        return send[:n_items].clone()
    recv = torch.empty(world * per_rank, dtype=torch.float32, device=local_scores.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv[:n_items]


class PinnedFeeder:
    """Double-buffered pinned host -> device staging of (B,N) fp32 batches on a side stream."""

    def __init__(self, batch_size, n_samples, device):
        self.device = torch.device(device)
        self.host = [torch.empty(batch_size, n_samples, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.dev = [torch.empty(batch_size, n_samples, dtype=torch.float32, device=self.device) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(self.device)
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        self.slot = 0

    def stage(self, fill_fn, rows):
        """fill_fn(host_view) writes `rows` utterances into pinned memory; returns (slot, rows)."""
        s = self.slot
        self.slot ^= 1
        self.free[s].synchronize()          # the forward that consumed this slot has finished
        fill_fn(self.host[s][:rows])
        with torch.cuda.stream(self.copy_stream):
            self.dev[s][:rows].copy_(self.host[s][:rows], non_blocking=True)
            self.ready[s].record(self.copy_stream)
        return s, rows

    def take(self, staged):
        s, rows = staged
        torch.cuda.current_stream(self.device).wait_event(self.ready[s])
        return self.dev[s][:rows]

    def release(self, staged):
        self.free[staged[0]].record(torch.cuda.current_stream(self.device))


def score_utterances(model, n_items, load_batch, n_samples, batch_size, device, rank=0, world=1, group=None,
                     preemph=False, coef=0.97):
    """Score items [0, n_items) sharded over `world` ranks; returns all scores (fp32, global order) on every rank.

    load_batch(lo, hi, out): fills the pinned host tensor `out` (hi-lo, n_samples) with utterances lo..hi-1
    (the reference's DataLoader role, main.py:200-209).  `model` is one of this package's model classes in
    eval mode on `device`.
    """
    lo, hi, per = shard_range(n_items, rank, world)
    feeder = PinnedFeeder(batch_size, n_samples, device)
    local = torch.empty(max(hi - lo, 0), dtype=torch.float32, device=device)
    ranges = batch_ranges(lo, hi, batch_size)
    eng = model.engine()
    staged = feeder.stage(lambda out, r=ranges[0]: load_batch(r[0], r[1], out), ranges[0][1] - ranges[0][0]) if ranges else None
    for i, (b_lo, b_hi) in enumerate(ranges):
        cur = staged
        if i + 1 < len(ranges):
            nxt = ranges[i + 1]
            staged = feeder.stage(lambda out, r=nxt: load_batch(r[0], r[1], out), nxt[1] - nxt[0])
        x = feeder.take(cur)
        logits = eng.forward(x, preemph=preemph, coef=coef)          # main.py:210
        local[b_lo - lo: b_hi - lo] = logits[:, 1]                    # main.py:212 (kept on device)
        feeder.release(cur)
    return gather_scores(local, n_items, per, group)


def write_score_file(path, utt_ids, scores):
    """'{utt} {score}' per line (reference main.py:216-219)."""
    scores = scores.detach().cpu().tolist()
    with open(path, "a+") as fh:
        for u, s in zip(utt_ids, scores):
            fh.write("{} {}\n".format(u, s))
