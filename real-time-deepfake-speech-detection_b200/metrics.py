"""Score sink and evaluation metrics: the step *after* the scoring path.

The reference pulls every batch back to the host -- ``batch_x[:, 1].data.cpu().numpy()`` in
``produce_evaluation_file`` (main.py:210-214), ``loss.item()`` and ``.sum().item()`` in ``Trainer._test``
(trainer.py:108-113) -- which serialises the GPU behind a D2H round trip per batch.  ``ScoreSink`` keeps the
bona-fide scores and the loss / accuracy accumulators on the device (one ``score_sink`` kernel per batch) and
copies them out once, at ``finish``.  ``equal_error_rate`` computes the integer ROC on the device
(``roc_counts``: TP/FP at every threshold, bit-exact) and solves the EER crossing in closed form, which is what
``brentq(lambda x: 1 - x - interp1d(fpr, tpr)(x), 0, 1)`` converges to (trainer.py:134-139).
"""
import ctypes

import torch

from .rtdf_runtime import native


class ScoreSink:
    def __init__(self, capacity, device, class_weight=None):
        self.lib = native.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ScoreSink lives on a CUDA device (no CPU path)")
        self.scores = torch.full((max(int(capacity), 1),), float("nan"), dtype=torch.float32, device=self.device)
        self.acc = torch.zeros(3, dtype=torch.float64, device=self.device)
        self.weight = None if class_weight is None else torch.as_tensor(class_weight, dtype=torch.float32).to(self.device)
        self.count = 0

    def push(self, logits, labels=None):
        """logits (B,2) fp32 CUDA; labels optional int64 (B,) CUDA (1 = bona fide).  No host synchronisation."""
        B = logits.shape[0]
        if B == 0:
            return
        if not logits.is_cuda or logits.dtype != torch.float32 or logits.shape[1] != 2:
            raise ValueError("ScoreSink.push expects (B,2) fp32 CUDA logits")
        if self.count + B > self.scores.numel():
            raise ValueError("ScoreSink capacity exceeded")
        logits = logits.contiguous()
        if labels is not None:
            labels = labels.to(device=self.device, dtype=torch.int64).contiguous().view(-1)
        dst = self.scores[self.count:self.count + B]
        with torch.cuda.device(self.device):
            native.check(self.lib.rtdf_score_sink(
                native.ptr(logits), B, native.ptr(labels), native.ptr(self.weight), native.ptr(dst),
                native.ptr(self.acc) if labels is not None else None,
                torch.cuda.current_stream(self.device).cuda_stream), "rtdf_score_sink")
        self.count += B

    def device_scores(self):
        return self.scores[:self.count]

    def finish(self):
        """One D2H: (scores fp32 CPU tensor, eval_loss, accuracy_percent) -- the last two as Trainer._test returns
        them (trainer.py:122-131), or None when no labels were pushed."""
        scores = self.scores[:self.count].cpu()
        acc = self.acc.cpu().tolist()
        if acc[2] == 0:
            return scores, None, None
        return scores, acc[0] / acc[2], acc[1] / acc[2] * 100.0


def roc_counts(scores, labels):
    """(tp, fp) int32 device tensors: positives / negatives with score >= scores[i] (NaN scores -> -1)."""
    if not scores.is_cuda:
        raise RuntimeError("roc_counts: CUDA tensors only (no CPU path)")
    scores = scores.to(torch.float32).contiguous().view(-1)
    labels = labels.to(device=scores.device, dtype=torch.int64).contiguous().view(-1)
    n = scores.numel()
    tp = torch.empty(n, dtype=torch.int32, device=scores.device)
    fp = torch.empty(n, dtype=torch.int32, device=scores.device)
    with torch.cuda.device(scores.device):
        native.check(native.load().rtdf_roc_counts(native.ptr(scores), native.ptr(labels), n, native.ptr(tp),
                                                   native.ptr(fp), torch.cuda.current_stream(scores.device).cuda_stream),
                     "rtdf_roc_counts")
    return tp, fp


def eer_from_bracket(tp_a, fp_a, tp_b, fp_b, n_pos, n_neg):
    """EER (percent) on the ROC segment a -> b, a the last point with 1 - fpr - tpr >= 0, b the first past it."""
    fa, ta = fp_a / n_neg, tp_a / n_pos
    fb, tb = fp_b / n_neg, tp_b / n_pos
    g = 1.0 - fa - ta
    step = (fb - fa) + (tb - ta)
    t = g / step if step > 0 else 0.0
    return (fa + t * (fb - fa)) * 100.0


def equal_error_rate(scores, labels):
    """EER in percent of CUDA scores (higher = bona fide) against labels (1 = bona fide); NaN scores are ignored."""
    scores = scores.to(torch.float32).contiguous().view(-1)
    labels = labels.to(device=scores.device, dtype=torch.int64).contiguous().view(-1)
    tp, fp = roc_counts(scores, labels)
    valid = ~torch.isnan(scores)
    n_pos = int((labels[valid] != 0).sum().item())
    n_neg = int(valid.sum().item()) - n_pos
    if n_pos == 0 or n_neg == 0:
        raise ValueError("equal_error_rate needs both classes")
    keys = torch.empty(2, dtype=torch.int64, device=scores.device)
    with torch.cuda.device(scores.device):
        native.check(native.load().rtdf_roc_crossing(native.ptr(tp), native.ptr(fp), scores.numel(), n_pos, n_neg,
                                                     native.ptr(keys), torch.cuda.current_stream(scores.device).cuda_stream),
                     "rtdf_roc_crossing")
    ka, kb = (k & 0xFFFFFFFFFFFFFFFF for k in keys.cpu().tolist())
    tp_a, fp_a = ka & 0xFFFFFFFF, (ka >> 32) - (ka & 0xFFFFFFFF)
    tp_b, fp_b = kb & 0xFFFFFFFF, (kb >> 32) - (kb & 0xFFFFFFFF)
    return eer_from_bracket(tp_a, fp_a, tp_b, fp_b, n_pos, n_neg)
