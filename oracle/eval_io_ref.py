"""CPU restatement of the two ends of the scoring path (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Pinned against the reference's own functions executed unmodified by ``oracle/check_eval_io_against_reference.py``
(golden vectors in ``tests/golden/eval_io_*.npz``).

  * ``adjust_duration`` / ``adjust_duration_random_start`` -- reference data/test_set.py:201-248
  * ``f_state_dict_wrapper``                               -- reference utils.py:13-43
  * ``eval_loss_accuracy``                                 -- reference trainer.py:85-132 (``Trainer._test``)
  * ``calculate_eer``                                      -- reference trainer.py:134-139
  * ``score_lines``                                        -- reference main.py:216-219
"""
import random
from collections import OrderedDict

import numpy as np
import torch


def _tile_to(x, duration):
    # data/test_set.py:217-227: repeat x duration // len times, append the residue
    x = x.squeeze() if x.dim() == 2 else x
    n = len(x)
    if n < duration:
        parts = [x for _ in range(duration // n)]
        residue = duration % n
        if residue > 0:
            parts.append(x[0:residue])
        x = torch.cat(parts, dim=0)
    return x


def adjust_duration(x, duration):
    """data/test_set.py:201-227: first `duration` samples of the (tile-repeated) utterance."""
    return _tile_to(x, duration)[0:duration]


def adjust_duration_random_start(x, duration, rng=random):
    """data/test_set.py:229-246: random window of `duration` samples (one rng.randint draw per call)."""
    x = _tile_to(x, duration)
    start = rng.randint(0, len(x) - duration)
    return x[start:start + duration]


def f_state_dict_wrapper(state_dict, data_parallel=False):
    """utils.py:13-43."""
    new = OrderedDict()
    for k, v in state_dict.items():
        if data_parallel:
            new[k if k.startswith("module") else "module." + k] = v
        else:
            new[k[7:] if k.startswith("module") else k] = v
    return new


def eval_loss_accuracy(batches, class_weight):
    """trainer.py:85-132 with loss_fn = nn.CrossEntropyLoss(weight) (main.py:106,122).
    batches: iterable of (logits (B,2) fp32, labels (B,) int64).  Returns (eval_loss, accuracy_percent)."""
    loss_fn = torch.nn.CrossEntropyLoss(torch.as_tensor(class_weight, dtype=torch.float32))
    num_correct, num_total, loss_sum = 0.0, 0.0, 0.0
    for x, label in batches:
        label = label.view(-1).type(torch.int64)
        bs = x.size(0)
        num_total += bs
        loss = loss_fn(x, label)
        _, pred = x.max(dim=1)
        num_correct += (pred == label).sum(dim=0).item()
        loss_sum += loss.item() * bs
    return loss_sum / num_total, (num_correct / num_total) * 100


def calculate_eer(scores, labels):
    """trainer.py:134-139, verbatim recipe: sklearn ROC + brentq on the interpolated curve; percent."""
    from scipy.interpolate import interp1d
    from scipy.optimize import brentq
    from sklearn import metrics
    fpr, tpr, _ = metrics.roc_curve(labels, scores, pos_label=1)
    eer = brentq(lambda x: 1. - x - interp1d(fpr, tpr)(x), 0., 1.)
    return eer * 100


def roc_counts(scores, labels):
    """Integer ROC: tp[i] / fp[i] = positives / negatives with score >= scores[i] (numpy, O(n log n))."""
    scores = np.asarray(scores, dtype=np.float32)
    labels = np.asarray(labels).astype(bool)
    pos = np.sort(scores[labels])
    neg = np.sort(scores[~labels])
    tp = len(pos) - np.searchsorted(pos, scores, side="left")
    fp = len(neg) - np.searchsorted(neg, scores, side="left")
    return tp.astype(np.int32), fp.astype(np.int32)


def score_lines(utt_ids, scores):
    """main.py:216-219: '{utt} {score}' per line."""
    return ["{} {}\n".format(f, cm) for f, cm in zip(utt_ids, scores)]
