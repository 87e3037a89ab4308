"""CPU: the two ends of the path -- oracle vs the golden vectors made from the
reference's own functions, and the host-side logic of the product (crop starts, EER bracket solve, hook firing,
checkpoint key handling).  No CUDA calls."""
import os
import random
import types

import numpy as np
import pytest
import torch

from tests.util import ROOT, pkg

GOLD = os.path.join(ROOT, "tests", "golden")


def _ragged(seed, lengths):
    g = torch.Generator().manual_seed(seed)
    return [0.1 * torch.randn(int(n), generator=g) for n in lengths]


def test_oracle_fit_duration_matches_reference_golden():
    from oracle import eval_io_ref as E
    z = np.load(os.path.join(GOLD, "eval_io_fit_duration.npz"))
    utts = _ragged(int(z["seed"]), z["lengths"])
    D = int(z["duration"])
    fit = torch.stack([E.adjust_duration(u, D) for u in utts])
    assert np.array_equal(fit.numpy(), z["fit"])
    random.seed(int(z["rand_seed"]))
    rnd = torch.stack([E.adjust_duration_random_start(u, D) for u in utts])
    assert np.array_equal(rnd.numpy(), z["fit_random"])


def test_crop_starts_follow_the_reference_random_draws():
    staging = pkg("staging")
    z = np.load(os.path.join(GOLD, "eval_io_fit_duration.npz"))
    utts = _ragged(int(z["seed"]), z["lengths"])
    D = int(z["duration"])
    random.seed(int(z["rand_seed"]))
    starts = staging.crop_starts([len(u) for u in utts], D, random_start=True)
    for u, st, want in zip(utts, starts, z["fit_random"]):
        idx = (st + np.arange(D)) % len(u)      # what the fit kernel computes
        assert np.array_equal(u.numpy()[idx], want)
    assert staging.crop_starts([5, 9000], D) == [0, 0]
    with pytest.raises(ValueError):
        staging.crop_starts([0], D)


def _bracket(tp, fp, P, Q):
    """numpy restatement of roc_crossing_kernel."""
    ok = tp >= 0
    tp, fp = tp[ok].astype(np.int64), fp[ok].astype(np.int64)
    le = fp * P + tp * Q <= P * Q
    key = ((tp + fp) << 32) | tp
    a = key[le].max() if le.any() else 0
    b = key[~le].min()
    return (a & 0xFFFFFFFF, (a >> 32) - (a & 0xFFFFFFFF), b & 0xFFFFFFFF, (b >> 32) - (b & 0xFFFFFFFF))


def test_eer_closed_form_equals_reference_brentq():
    from oracle import eval_io_ref as E
    metrics = pkg("metrics")
    z = np.load(os.path.join(GOLD, "eval_io_eer.npz"))
    assert bool(z["pinned"])
    scores, labels = z["scores"], z["labels"]
    assert abs(E.calculate_eer(scores, labels) - float(z["eer"])) < 1e-12
    tp, fp = E.roc_counts(scores, labels)
    assert np.array_equal(tp, z["tp"]) and np.array_equal(fp, z["fp"])
    P, Q = int(labels.sum()), int((1 - labels).sum())
    eer = metrics.eer_from_bracket(*_bracket(tp, fp, P, Q), P, Q)
    assert abs(eer - float(z["eer"])) < 1e-8, (eer, float(z["eer"]))
    # more shapes: separable, inverted, heavy ties, tiny
    rng = np.random.default_rng(3)
    for n, shift, rnd in [(50, 5.0, None), (400, -1.0, None), (1000, 0.7, 0), (7, 1.0, None), (2000, 0.3, 1)]:
        lab = (rng.random(n) < 0.4).astype(np.int64)
        lab[0], lab[1] = 0, 1
        sc = (rng.standard_normal(n) + shift * lab).astype(np.float32)
        if rnd is not None:
            sc = np.round(sc, rnd)
        want = E.calculate_eer(sc, lab)
        tp, fp = E.roc_counts(sc, lab)
        P, Q = int(lab.sum()), int((1 - lab).sum())
        got = metrics.eer_from_bracket(*_bracket(tp, fp, P, Q), P, Q)
        assert abs(got - want) < 1e-7, (n, shift, rnd, got, want)


def test_eval_loss_accuracy_oracle_self_consistency():
    from oracle import eval_io_ref as E
    g = torch.Generator().manual_seed(0)
    x = torch.randn(6, 2, generator=g)
    y = torch.tensor([0, 1, 1, 0, 1, 0])
    loss, acc = E.eval_loss_accuracy([(x[:4], y[:4]), (x[4:], y[4:])], [0.9, 0.1])
    w = torch.tensor([0.9, 0.1])
    nll = -(torch.log_softmax(x, 1)[torch.arange(6), y])
    want = (4 * (w[y[:4]] * nll[:4]).sum() / w[y[:4]].sum() + 2 * (w[y[4:]] * nll[4:]).sum() / w[y[4:]].sum()) / 6
    assert abs(loss - float(want)) < 1e-6
    assert acc == pytest.approx(100.0 * float((x.argmax(1) == y).float().mean()))


def test_state_dict_wrapper_and_prefixed_checkpoints_load():
    utils = pkg("rtdf_utils")
    xa = pkg("models.xlsr_aasist")
    from oracle import eval_io_ref as E
    sd = {"module.a.weight": 1, "b.bias": 2}
    for dp in (False, True):
        assert list(utils.f_state_dict_wrapper(sd, dp).items()) == list(E.f_state_dict_wrapper(sd, dp).items())
    # a fine-tuned reference checkpoint carries DataParallel's "module." prefix (main.py:391-395)
    m = xa.My_XLSR_AASIST("cpu", None, num_layers=1)
    ckpt = utils.f_state_dict_wrapper(m.state_dict(), data_parallel=True)
    assert all(k.startswith("module.") for k in ckpt)
    m2 = xa.My_XLSR_AASIST("cpu", None, num_layers=1)
    m2.load_state_dict(utils.f_state_dict_wrapper(ckpt, data_parallel=False), strict=True)
    torch.nn.DataParallel(m2).load_state_dict(ckpt, strict=True)
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k


def test_custom_order_copy_of_teacher_layers():
    """main_kd.py:121-141: student.load_state_dict(teacher, strict=False) then per-index layer copies."""
    xa = pkg("models.xlsr_aasist")
    torch.manual_seed(0)
    teacher = xa.My_XLSR_AASIST("cpu", None, num_layers=4)
    order = [3, 1]
    student = xa.My_XLSR_AASIST("cpu", None, num_layers=2, order="custom", custom_order=[0, 1])
    student.load_state_dict(teacher.state_dict(), strict=False)
    for index, value in enumerate(order):
        student.ssl_model.model.encoder.layers[index].load_state_dict(
            teacher.ssl_model.model.encoder.layers[value].state_dict(), strict=False)
    for index, value in enumerate(order):
        a = student.ssl_model.model.encoder.layers[index].fc1.weight
        b = teacher.ssl_model.model.encoder.layers[value].fc1.weight
        assert torch.equal(a, b)
    assert torch.equal(student.LL.weight, teacher.LL.weight)


def test_forward_hooks_fire_with_fairseq_shaped_io():
    hooks = pkg("models._hooks")
    xa = pkg("models.xlsr_aasist")
    model = xa.My_XLSR_AASIST("cpu", None, num_layers=2)
    B, T = 3, 5
    layers = torch.arange(3 * B * T * 1024, dtype=torch.float32).view(3, B, T, 1024)
    feats = torch.ones(B, T, 1024)
    calls = {}

    class FakeEngine:
        def forward(self, x, want_taps=False, layer_taps=False, **kw):
            calls["args"] = (want_taps, layer_taps)
            logits = torch.zeros(B, 2)
            return (logits, {"layers": layers, "feats": feats}) if want_taps else logits

    x = torch.zeros(B, 16000)
    assert hooks.forward_with_hooks(model, FakeEngine(), x).shape == (B, 2)
    assert calls["args"] == (False, False)          # no hooks: plain forward, no taps
    seen = {}
    model.ssl_model.model.encoder.layers[1].register_forward_hook(
        lambda m, inp, out: seen.update(l1=(inp, out)))
    model.ssl_model.register_forward_hook(lambda m, inp, out: seen.update(ssl=(inp, out)))
    hooks.forward_with_hooks(model, FakeEngine(), x)
    assert calls["args"] == (True, True)
    inp, out = seen["l1"]
    assert inp[0].shape == (T, B, 1024) and torch.equal(inp[0], layers[1].transpose(0, 1))
    assert torch.equal(out[0], layers[2].transpose(0, 1)) and out[1] == (None, None)
    assert seen["ssl"][0][0] is x and seen["ssl"][1] is feats


def test_new_host_modules_have_no_cpu_path():
    """staging / metrics / scoring pipeline refuse CPU devices and CPU tensors (the product path never falls back)."""
    staging, metrics, scoring = pkg("staging"), pkg("metrics"), pkg("scoring")
    with pytest.raises(RuntimeError):
        staging.UtteranceStager(4000, 4, "cpu")
    with pytest.raises(RuntimeError):
        metrics.ScoreSink(8, "cpu")
    with pytest.raises(RuntimeError):
        metrics.roc_counts(torch.zeros(4), torch.zeros(4, dtype=torch.int64))
    with pytest.raises(RuntimeError):
        staging.fit_duration(torch.zeros(10), torch.tensor([0, 10]), 4)
    with pytest.raises(RuntimeError):
        scoring.ScoringPipeline(object(), 8, 4, 16000, "cpu")


def test_score_file_format_matches_reference_lines(tmp_path):
    from oracle import eval_io_ref as E
    scoring = pkg("scoring")
    ids = ["LA_E_1000147", "LA_E_1000273", "DF_E_2000011"]
    scores = torch.tensor([1.5, -0.25, 3.0e-5])
    path = tmp_path / "scores.txt"
    scoring.write_score_file(str(path), ids, scores)
    assert path.read_text().splitlines(keepends=True) == E.score_lines(ids, scores.tolist())
