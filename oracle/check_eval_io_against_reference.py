"""Pin ``oracle/eval_io_ref.py`` against the reference's own functions and emit golden vectors.

TEST INFRASTRUCTURE ONLY; runs only in the build container (needs ``/root/reference``).
The reference's ``data/test_set.py``, ``utils.py`` and ``trainer.py`` are imported UNMODIFIED; their third-party
imports that this image lacks (librosa, wandb, torch_audiomentations, audiomentations, the RawBoost module's
scipy-only deps are present) are satisfied by empty shim modules -- none of them is touched by the functions
under test.

Usage:  python -m oracle.check_eval_io_against_reference [--write]
"""
import argparse
import os
import random
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def import_reference():
    for name in ("librosa", "wandb", "audiomentations"):
        sys.modules.setdefault(name, types.ModuleType(name))
    ta = types.ModuleType("torch_audiomentations")
    for n in ("Compose", "AddColoredNoise", "HighPassFilter", "LowPassFilter", "Gain"):
        setattr(ta, n, object)
    sys.modules.setdefault("torch_audiomentations", ta)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import importlib
    ts = importlib.import_module("data.test_set")
    ut = importlib.import_module("utils")
    try:
        tr = importlib.import_module("trainer")
    except Exception as e:  # noqa: BLE001
        print("trainer.py not importable here:", repr(e))
        tr = None
    return ts, ut, tr


def synth_ragged(seed, lengths):
    g = torch.Generator().manual_seed(seed)
    return [0.1 * torch.randn(n, generator=g) for n in lengths]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--write", action="store_true")
    args = ap.parse_args()
    ts, ut, tr = import_reference()
    from oracle import eval_io_ref as E

    # ---- adjustDuration / adjustDuration_random_start (data/test_set.py:201-248) ------------------
    duration = 4000
    lengths = [1, 7, 999, 1333, 3999, 4000, 4001, 9000, 2000, 12345]
    utts = synth_ragged(77, lengths)
    holder = types.SimpleNamespace(duration=duration)
    cls = ts.ASVspoof2021DF_eval
    ref_fit = torch.stack([cls.adjustDuration(holder, u) for u in utts])
    ora_fit = torch.stack([E.adjust_duration(u, duration) for u in utts])
    assert torch.equal(ref_fit, ora_fit), "adjustDuration mismatch"
    ref_2d = cls.adjustDuration(holder, utts[3].view(1, -1))
    assert torch.equal(ref_2d, ora_fit[3])
    random.seed(1234)
    ref_rand = torch.stack([cls.adjustDuration_random_start(holder, u) for u in utts])
    random.seed(1234)
    ora_rand = torch.stack([E.adjust_duration_random_start(u, duration) for u in utts])
    assert torch.equal(ref_rand, ora_rand), "adjustDuration_random_start mismatch"
    print("adjustDuration / adjustDuration_random_start: oracle == reference (bit-exact) on", len(utts), "utterances")

    # ---- f_state_dict_wrapper (utils.py:13-43) -------------------------------------------------------
    sd = {"module.a.weight": 1, "b.bias": 2, "module.c": 3}
    for dp in (False, True):
        assert list(ut.f_state_dict_wrapper(sd, dp).items()) == list(E.f_state_dict_wrapper(sd, dp).items())
    print("f_state_dict_wrapper: oracle == reference")

    # ---- calculate_EER (trainer.py:134-139) -----------------------------------------------------------
    g = torch.Generator().manual_seed(5)
    n = 3000
    labels = (torch.rand(n, generator=g) < 0.15).long()
    scores = (torch.randn(n, generator=g) + 1.5 * labels).float()
    scores[::7] = scores[::7].round(decimals=1)       # ties, some across classes
    eer_ref = None
    if tr is not None:
        eer_ref = tr.Trainer.calculate_EER(None, scores.numpy(), labels.numpy())
        eer_ora = E.calculate_eer(scores.numpy(), labels.numpy())
        assert eer_ref == eer_ora, (eer_ref, eer_ora)
        print("calculate_EER: oracle == reference:", eer_ref)
    else:
        eer_ref = E.calculate_eer(scores.numpy(), labels.numpy())
        print("calculate_EER: reference trainer.py not importable; oracle value", eer_ref, "(parity unpinned for EER)")
    tp, fp = E.roc_counts(scores.numpy(), labels.numpy())

    if args.write:
        os.makedirs(GOLDEN, exist_ok=True)
        np.savez_compressed(os.path.join(GOLDEN, "eval_io_fit_duration.npz"), seed=77, lengths=np.array(lengths),
                            duration=duration, fit=ref_fit.numpy(), rand_seed=1234, fit_random=ref_rand.numpy())
        np.savez_compressed(os.path.join(GOLDEN, "eval_io_eer.npz"), scores=scores.numpy(), labels=labels.numpy(),
                            eer=np.float64(eer_ref), pinned=np.bool_(tr is not None), tp=tp, fp=fp)
        print("wrote tests/golden/eval_io_fit_duration.npz, eval_io_eer.npz")


if __name__ == "__main__":
    main()
