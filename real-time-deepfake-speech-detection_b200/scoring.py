"""Data-parallel scoring driver: utterances sharded across the GPUs of one box, one collective.

Mirrors what the reference's scoring callers do per batch -- ``produce_evaluation_file``
(reference main.py:199-221) and ``Trainer._test`` (trainer.py:85-132): ``model(batch_x)`` under
``no_grad`` and ``score = out[:, 1]`` -- but

  * rank r of W scores the contiguous index range [r*S, min((r+1)*S, n)), S = ceil(n/W)
    (SURVEY.md section 8e); every rank holds a full weight replica, there is no data-path collective;
  * scores stay on the device and are exchanged with ONE all-gather of fp32[S] per rank
    (NCCL over NVLink on GPUs; gloo in the CPU tests of the host logic), tail padded with NaN;
  * H2D copies come from pinned host memory on a side stream, overlapped with the previous forward; the scores
    of every batch are copied back asynchronously (``ScoringPipeline``).

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


class ScoringPipeline:
    """The scoring loop of ``produce_evaluation_file`` (reference main.py:199-221) as a three-stage pipeline:

      copy stream     H2D of batch i+1 from pinned host memory        (main.py:209  batch_x.to(device))
      compute stream  forward of batch i (one CUDA-graph replay)       (main.py:210  model(batch_x))
      compute stream  scores of batch i -> device vector -> async D2H  (main.py:212  batch_x[:, 1].data.cpu())

    The reference serialises the three per batch (pageable copy, forward, blocking ``.cpu()``); here the host never
    waits for the device inside the loop except to recycle one of `depth` input slots, and the only full
    synchronisation is in ``finish``.  Every batch still crosses PCIe in both directions.
    """

    def __init__(self, model, capacity, batch_size, n_samples, device, preemph=False, coef=0.97, depth=2,
                 regime="throughput"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ScoringPipeline runs on CUDA devices only (no CPU path)")
        self.eng = model.engine()
        self.preemph, self.coef = preemph, coef
        # large-batch kernels only: a score must not depend on the batch (ragged tail) or the shard it lands in
        self.regime = regime
        self.batch_size, self.n_samples = int(batch_size), int(n_samples)
        self.dev_in = [torch.empty(batch_size, n_samples, dtype=torch.float32, device=self.device) for _ in range(depth)]
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.free = [torch.cuda.Event() for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(self.device)
        self.scores_dev = torch.full((max(int(capacity), 1),), PAD, dtype=torch.float32, device=self.device)
        self.scores_host = torch.full((max(int(capacity), 1),), PAD, dtype=torch.float32).pin_memory()
        self.count = 0
        self.step = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def push(self, host_batch):
        """host_batch: (b, n_samples) fp32 CPU tensor, b <= batch_size; pinned memory makes the copy asynchronous.
        A CUDA tensor is accepted too and scored in place (no H2D)."""
        b = host_batch.shape[0]
        if b == 0:
            return
        if b > self.batch_size or host_batch.shape[1] != self.n_samples:
            raise ValueError(f"batch shape {tuple(host_batch.shape)} does not fit ({self.batch_size}, {self.n_samples})")
        if self.count + b > self.scores_dev.numel():
            raise ValueError("ScoringPipeline capacity exceeded")
        cur = torch.cuda.current_stream(self.device)
        if host_batch.is_cuda:
            # already resident (e.g. synthesised on the device): no staging copy, plain stream order
            x = host_batch.to(torch.float32).contiguous()
            logits = self.eng.forward(x, preemph=self.preemph, coef=self.coef, regime=self.regime)
        else:
            slot = self.step % len(self.dev_in)
            self.step += 1
            self.free[slot].synchronize()                    # the forward that read this slot has finished
            x = self.dev_in[slot][:b]
            with torch.cuda.stream(self.copy_stream):
                x.copy_(host_batch, non_blocking=True)
                self.ready[slot].record(self.copy_stream)
            cur.wait_event(self.ready[slot])
            logits = self.eng.forward(x, preemph=self.preemph, coef=self.coef, regime=self.regime)
            self.free[slot].record(cur)
            self.h2d_bytes += host_batch.numel() * 4
        dst = self.scores_dev[self.count:self.count + b]
        dst.copy_(logits[:, 1])                                                          # stays on the device ...
        self.scores_host[self.count:self.count + b].copy_(dst, non_blocking=True)       # ... and goes home asynchronously
        self.count += b
        self.d2h_bytes += b * 4

    def device_scores(self):
        return self.scores_dev[:self.count]

    def finish(self):
        """Synchronise once; returns the scores pushed so far as a CPU tensor (a view of pinned memory)."""
        torch.cuda.current_stream(self.device).synchronize()
        return self.scores_host[:self.count]


def score_utterances(model, n_items, load_batch, n_samples, batch_size, device, rank=0, world=1, group=None,
                     preemph=False, coef=0.97, zero_copy=False):
    """Score items [0, n_items) sharded over `world` ranks; returns all scores (fp32, global order) on every rank.

    load_batch(lo, hi, out): fills the pinned host tensor `out` (hi-lo, n_samples) with utterances lo..hi-1
    (the reference's DataLoader role, main.py:200-209) and returns None -- or returns its own (hi-lo, n_samples) fp32
    CPU tensor (ideally pinned, e.g. a slice of a resident pool), which is then copied to the device instead of `out`
    and must stay untouched until this call returns (zero_copy=True: load_batch is called with out=None and MUST return
    the batch -- no pinned staging buffers are allocated at all).  `model` is one of this package's model classes in eval mode on
    `device`.  The forwards run in the throughput regime, so a score does not depend on the batch or shard an utterance
    lands in: the returned vector is bit-identical for every `world` / `batch_size`.
    """
    lo, hi, per = shard_range(n_items, rank, world)
    pipe = ScoringPipeline(model, max(hi - lo, 1), batch_size, n_samples, device, preemph=preemph, coef=coef, depth=2)
    host = [None, None, None]     # pinned staging buffers, allocated on first use (a zero-copy load_batch never needs them)

    for i, (b_lo, b_hi) in enumerate(batch_ranges(lo, hi, batch_size)):
        # 3 host buffers for 2 device slots: push(i) waits for forward(i-2), whose H2D (the last reader of buffer
        # (i-2) % 3 ... and of buffer i % 3 = (i-3) % 3) is then complete
        if zero_copy:
            own = load_batch(b_lo, b_hi, None)
            if own is None:
                raise ValueError("score_utterances(zero_copy=True): load_batch must return the batch tensor")
        else:
            if host[i % 3] is None:
                host[i % 3] = torch.empty(batch_size, n_samples, dtype=torch.float32).pin_memory()
            out = host[i % 3][: b_hi - b_lo]
            own = load_batch(b_lo, b_hi, out)
            if own is None:
                own = out
        pipe.push(own)
    torch.cuda.current_stream(torch.device(device)).synchronize()
    return gather_scores(pipe.device_scores(), n_items, per, group)


def write_score_file(path, utt_ids, scores):
    """'{utt} {score}' per line (reference main.py:216-219)."""
    scores = scores.detach().cpu().tolist()
    with open(path, "a+") as fh:
        for u, s in zip(utt_ids, scores):
            fh.write("{} {}\n".format(u, s))
