"""GPU parity of the two ends of the path against the oracle and the golden vectors
generated from the reference's own functions.  Integer / copy work is compared bit-exactly."""
import os
import random

import numpy as np
import pytest
import torch

from tests.util import ROOT, build_pair, pkg

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


def _ragged(seed, lengths):
    g = torch.Generator().manual_seed(seed)
    return [0.1 * torch.randn(int(n), generator=g) for n in lengths]


def test_fit_duration_bit_exact_vs_reference_golden():
    staging = pkg("staging")
    z = np.load(os.path.join(GOLD, "eval_io_fit_duration.npz"))
    utts = _ragged(int(z["seed"]), z["lengths"])
    D = int(z["duration"])
    packed = torch.cat(utts).cuda()
    offs = torch.tensor(np.concatenate([[0], np.cumsum(z["lengths"])]), dtype=torch.int64).cuda()
    out = staging.fit_duration(packed, offs, D)
    assert np.array_equal(out.cpu().numpy(), z["fit"])
    random.seed(int(z["rand_seed"]))
    starts = staging.crop_starts([len(u) for u in utts], D, random_start=True)
    out = staging.fit_duration(packed, offs, D, starts=torch.tensor(starts, dtype=torch.int32).cuda())
    assert np.array_equal(out.cpu().numpy(), z["fit_random"])


@pytest.mark.parametrize("D", [4000, 4001, 1, 2])
def test_fit_duration_fused_preemphasis(D):
    from oracle import eval_io_ref as E
    from oracle import models_ref as O
    staging = pkg("staging")
    lengths = [1, 2, 3, 777, D, D + 5, 3 * D + 1]
    utts = _ragged(5, lengths)
    packed = torch.cat(utts).cuda()
    offs = torch.tensor(np.concatenate([[0], np.cumsum(lengths)]), dtype=torch.int64).cuda()
    got = staging.fit_duration(packed, offs, D, preemph=0.97).cpu()
    fit = torch.stack([E.adjust_duration(u, D) for u in utts])
    if D > 1:
        want = O.pre_emphasis(fit, 0.97).view(len(utts), D)
        assert (got - want).abs().max() <= 1e-7      # fp32 x - 0.97*x_prev: fma vs mul+add rounding
    plain = staging.fit_duration(packed, offs, D).cpu()
    assert torch.equal(plain, fit)


def test_stager_double_buffering_matches_oracle():
    from oracle import eval_io_ref as E
    staging = pkg("staging")
    D, B = 3000, 4
    st = staging.UtteranceStager(D, B, "cuda", random_start=True, rng=random.Random(9))
    rng_ref = random.Random(9)
    # lengths straddle the duration: shorter ones are tiled on the device, longer ones cropped at a random start on the host
    batches = [_ragged(100 + i, [50 + 1400 * j + 7 * i for j in range(B if i != 4 else 2)]) for i in range(5)]
    tickets = [st.stage(batches[0])]
    for i, batch in enumerate(batches):
        if i + 1 < len(batches):
            tickets.append(st.stage(batches[i + 1]))
        got = st.take(tickets[i]).clone()
        st.release(tickets[i])
        want = torch.stack([E.adjust_duration_random_start(u, D, rng_ref) for u in batch])
        assert torch.equal(got.cpu(), want), i
    assert st.h2d_bytes < sum(len(b) for b in batches) * D * 4 + 4096     # never more than `duration` samples per utterance
    with pytest.raises(ValueError):
        st.stage([])
    st2 = staging.UtteranceStager(D, B, "cuda")
    st2.stage(batches[0])
    st2.stage(batches[1])
    with pytest.raises(RuntimeError, match="outstanding"):
        st2.stage(batches[2])                      # both slots still hold unreleased tickets


def test_score_sink_matches_trainer_test_accumulators():
    from oracle import eval_io_ref as E
    metrics = pkg("metrics")
    g = torch.Generator().manual_seed(11)
    sizes = [64, 64, 7, 1, 300]
    batches = [(3 * torch.randn(n, 2, generator=g), (torch.rand(n, generator=g) < 0.3).long()) for n in sizes]
    batches[2][0][0] = torch.tensor([0.5, 0.5])       # tie -> class 0 (torch.max picks the first index)
    sink = metrics.ScoreSink(sum(sizes), "cuda", class_weight=[0.9, 0.1])
    for x, y in batches:
        sink.push(x.cuda(), y.cuda())
    scores, loss, acc = sink.finish()
    want_loss, want_acc = E.eval_loss_accuracy(batches, [0.9, 0.1])
    assert torch.equal(scores, torch.cat([x[:, 1] for x, _ in batches]))      # bit-exact copy of out[:,1]
    assert abs(loss - want_loss) < 2e-6 * max(1.0, abs(want_loss))
    assert acc == pytest.approx(want_acc, abs=1e-9)
    plain = metrics.ScoreSink(10, "cuda")
    plain.push(batches[3][0].cuda())
    s, l, a = plain.finish()
    assert l is None and a is None and torch.equal(s, batches[3][0][:, 1])
    with pytest.raises(ValueError):
        plain.push(torch.zeros(64, 2, device="cuda"))


def test_roc_counts_bit_exact_and_eer_vs_reference_golden():
    from oracle import eval_io_ref as E
    metrics = pkg("metrics")
    z = np.load(os.path.join(GOLD, "eval_io_eer.npz"))
    scores, labels = torch.tensor(z["scores"]).cuda(), torch.tensor(z["labels"]).cuda()
    tp, fp = metrics.roc_counts(scores, labels)
    assert np.array_equal(tp.cpu().numpy(), z["tp"]) and np.array_equal(fp.cpu().numpy(), z["fp"])
    assert abs(metrics.equal_error_rate(scores, labels) - float(z["eer"])) < 1e-8
    # NaN padding (score gather of a ragged last shard) is ignored
    pad = torch.cat([scores, torch.full((5,), float("nan"), device="cuda")])
    lab = torch.cat([labels, torch.zeros(5, dtype=torch.int64, device="cuda")])
    assert abs(metrics.equal_error_rate(pad, lab) - float(z["eer"])) < 1e-8
    # a larger random case with ties against the oracle (sklearn + brentq, as trainer.py:134-139)
    g = torch.Generator().manual_seed(1)
    n = 20011
    lab = (torch.rand(n, generator=g) < 0.1).long()
    sc = (torch.randn(n, generator=g) + 2.0 * lab).float().round(decimals=2)
    tp, fp = metrics.roc_counts(sc.cuda(), lab.cuda())
    wtp, wfp = E.roc_counts(sc.numpy(), lab.numpy())
    assert np.array_equal(tp.cpu().numpy(), wtp) and np.array_equal(fp.cpu().numpy(), wfp)
    assert abs(metrics.equal_error_rate(sc.cuda(), lab.cuda()) - E.calculate_eer(sc.numpy(), lab.numpy())) < 1e-7


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 0.05)])
def test_layer_taps_and_forward_hooks(precision, tol):
    """f3: the residual stream entering / leaving every transformer layer equals the oracle's layer I/O, and
    forward hooks registered on encoder.layers[i] (the KD hook points) receive it in fairseq's (T,B,C) layout."""
    from oracle import models_ref as O
    ora, prod = build_pair("My_XLSR_AASIST", precision, num_layers=3, order="first")
    x = O.synth_waveforms(2, 16000, seed=2021)
    want = {}
    hs = [layer.register_forward_hook(lambda m, i, o, k=k: want.update({k: (i[0], o)}))
          for k, layer in enumerate(ora.ssl_model.model.encoder.layers)]
    with torch.no_grad():
        ref = ora(x)
    for h in hs:
        h.remove()
    seen = {}
    layers = prod.ssl_model.model.encoder.layers
    layers[0].register_forward_hook(lambda m, i, o: seen.update({0: (i[0], o[0])}))
    layers[2].register_forward_hook(lambda m, i, o: seen.update({2: (i[0], o[0])}))
    with torch.no_grad():
        got = prod(x.cuda())
    assert (got.cpu() - ref).abs().max() <= (1e-4 if precision == "fp32" else 1e-2)
    for k in (0, 2):
        xin, xout = seen[k]
        assert xin.shape == (49, 2, 1024)
        scale = want[k][1].abs().max()
        assert (xin.transpose(0, 1).cpu() - want[k][0]).abs().max() <= tol * scale
        assert (xout.transpose(0, 1).cpu() - want[k][1]).abs().max() <= tol * scale


def test_scoring_pipeline_matches_direct_forward_bit_exactly():
    """H2D / forward / D2H pipelining must not change a single score (ragged last batch included)."""
    from oracle import models_ref as O
    scoring = pkg("scoring")
    _, prod = build_pair("My_XLSR_AASIST", "bf16", num_layers=1, order="first")
    N, B = 16000, 4
    x = O.synth_waveforms(11, N, seed=5)
    eng = prod.engine()
    with torch.no_grad():
        # the pipeline pins the throughput regime (batch-composition-invariant kernels, include/rtdf.h rtdf_regime)
        want = torch.cat([eng.forward(x[i:i + B].cuda(), regime="throughput")[:, 1].cpu() for i in range(0, 11, B)])
        auto = torch.cat([prod(x[i:i + B].cuda())[:, 1].cpu() for i in range(0, 11, B)])   # model(x): streaming kernels
    assert (auto - want).abs().max() <= 1e-3          # bf16: the two regimes differ at rounding level only
    one_batch = eng.forward(x.cuda(), regime="throughput")[:, 1].cpu()
    assert torch.equal(one_batch, want)               # ... and the throughput regime does not see the batching at all
    pipe = scoring.ScoringPipeline(prod, 11, B, N, "cuda")
    for i in range(0, 11, B):
        pipe.push(x[i:i + B].pin_memory())
    got = pipe.finish()
    assert torch.equal(got, want)
    assert pipe.h2d_bytes == 11 * N * 4 and pipe.d2h_bytes == 11 * 4
    with pytest.raises(ValueError):
        pipe.push(x[:1].pin_memory())            # capacity exceeded

    def load(lo, hi, out):
        out.copy_(x[lo:hi])
    all_scores = scoring.score_utterances(prod, 11, load, N, B, "cuda")
    assert torch.equal(all_scores.cpu(), want)
    on_dev = scoring.score_utterances(prod, 11, lambda lo, hi, out: x[lo:hi].cuda(), N, 5, "cuda", zero_copy=True)  # resident batches
    assert torch.equal(on_dev.cpu(), want)
    pool = x.pin_memory()
    zero_copy = scoring.score_utterances(prod, 11, lambda lo, hi, out: pool[lo:hi], N, 3, "cuda", zero_copy=True)   # other batch size
    assert torch.equal(zero_copy.cpu(), want)


def test_streaming_static_buffers_match_forward():
    """Engine.static_input / forward_static (zero-copy streaming calls) give the scores of Engine.forward bit for bit."""
    from oracle import models_ref as O
    _, prod = build_pair("My_XLSR_AASIST", "bf16", num_layers=1, order="first")
    eng = prod.engine()
    x = O.synth_waveforms(3, 16000, seed=9)
    want = [eng.forward(x[i:i + 1].cuda()).cpu() for i in range(3)]
    buf = eng.static_input(1, 16000)
    host = x.pin_memory()
    for i in range(3):
        buf.copy_(host[i:i + 1], non_blocking=True)
        got = eng.forward_static(1, 16000).cpu()
        assert torch.equal(got, want[i]), i
    assert torch.equal(eng.forward(buf).cpu(), want[2])       # the static buffer itself is accepted without a copy


def test_engine_cache_invalidation_and_regime_argument():
    """engine_for re-packs when a parameter's version counter moves; a write through .data is invisible to it until
    rtdf_runtime.invalidate(model); unknown regimes are rejected."""
    _, prod = build_pair("My_XLSR_AASIST", "bf16", num_layers=1, order="first")
    rt = pkg("rtdf_runtime")
    from oracle import models_ref as O
    x = O.synth_waveforms(2, 16000, seed=21).cuda()
    y0 = prod(x).clone()
    eng0 = prod.engine()
    with torch.no_grad():
        prod.out_layer.bias.add_(0.5)                      # in-place op: version counter moves, engine rebuilt
    assert prod.engine() is not eng0
    y1 = prod(x)
    assert float((y1 - y0 - 0.5).abs().max()) < 1e-5
    prod.out_layer.bias.data.add_(0.25)                    # .data write: not seen ...
    assert float((prod(x) - y1).abs().max()) == 0.0
    rt.invalidate(prod)                                    # ... until the cached engine is dropped
    assert float((prod(x) - y1 - 0.25).abs().max()) < 1e-5
    with pytest.raises(ValueError):
        prod.engine().forward(x, regime="fastest")
