"""Run the reference's own scoring callers, UNMODIFIED, against this package's ``models`` / ``data.preprocess``.

Executed as a script in a fresh interpreter by tests/test_dropin_*.py (so that the top-level names ``models``,
``data``, ``utils``, ``main``, ``trainer`` of this experiment never leak into the pytest process):

    python tests/dropin_runner.py --mode stub|cuda --out result.json

sys.path order is the one INTEGRATION.md documents: <this package's directory> ahead of <reference root>.  The
reference root is /root/reference when it exists (build container), else the git-ignored archive
oracle/_ref/reference_py.zip that ``__graft_entry__.build()`` made (GPU box).

What runs:
  * ``main.produce_evaluation_file(dataset, model, device, save_path, batch_size)``   (reference main.py:199-221)
  * ``trainer.Trainer._test(loader)`` with the mirror ``data.preprocess.PreEmphasis``  (reference trainer.py:85-132)

mode=stub: CPU box; the model is this package's My_XLSR_AASIST whose engine is replaced by a stub returning
           deterministic logits -- checks the wiring (imports, class lookup, forward contract, score file, loss/acc).
mode=cuda: 1-layer student on cuda:0 through librtdf.so; scores are compared with the CPU oracle.
"""
import argparse
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG_DIR = os.path.join(ROOT, "real-time-deepfake-speech-detection_b200")


class _Cfg:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["stub", "cuda"], required=True)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()

    sys.path.insert(0, ROOT)
    from oracle import build_ref
    ref_root = build_ref.reference_root()
    if ref_root is None:
        json.dump({"skipped": "no reference tree or archive"}, open(args.out, "w"))
        return
    build_ref.install_third_party_stubs()
    sys.path[:0] = [PKG_DIR, ref_root]

    import torch
    import main as ref_main                        # the reference's main.py
    import main_kd as ref_main_kd                  # the reference's main_kd.py (student / teacher scoring entry point)
    import trainer as ref_trainer                  # the reference's trainer.py
    import models
    import utils
    import data.preprocess as prep

    res = {"ref_root": ref_root,
           "models_from": models.__file__, "preprocess_from": prep.__file__, "utils_from": utils.__file__,
           "main_from": ref_main.__file__, "trainer_from": ref_trainer.__file__}
    assert res["models_from"].startswith(PKG_DIR) and res["preprocess_from"].startswith(PKG_DIR), res
    assert res["utils_from"].startswith(ref_root) and res["main_from"].startswith(ref_root), res
    # class lookup by name in main's globals (reference main.py:76-84)
    model_class = vars(ref_main)["My_XLSR_AASIST"]
    assert model_class.__module__ == "models.xlsr_aasist", model_class.__module__
    assert vars(ref_main)["ConformerModel"].__module__ == "models.conformer_baseline"
    assert vars(ref_main_kd)["MyConformerModel"].__module__ == "models.conformer_baseline"      # main_kd.py:20-22
    assert vars(ref_main_kd)["My_XLSR_AASIST"].__module__ == "models.xlsr_aasist"
    assert utils.f_state_dict_wrapper({"module.a": 1})  # the reference's own helper is the one in use

    n_utt, n_samples, batch = 8, 16000, 3          # ragged last batch of 2 (drop_last=False); a last batch of 1 would hit
                                                   # the reference's own PreEmphasis squeeze() hazard (preprocess.py:27)
    g = torch.Generator().manual_seed(11)
    waves = 0.1 * torch.randn(n_utt, n_samples, generator=g)
    labels = torch.tensor([0, 1, 1, 0, 1, 0, 1, 1])

    class FakeSet(torch.utils.data.Dataset):       # Dataset.__getitem__ -> (utt_id, wav[N], label)  (test_set.py:178-199)
        def __len__(self):
            return n_utt

        def __getitem__(self, i):
            return f"utt_{i:03d}", waves[i], int(labels[i])

    device = "cuda:0" if args.mode == "cuda" else "cpu"
    torch.manual_seed(1024)
    model = model_class(device=device, ssl_cpkt_path=None, num_layers=1, order="first").to(device)
    expect = None
    if args.mode == "stub":
        class StubEngine:                          # stands in for rtdf_runtime.Engine on a box without a GPU
            calls = 0

            def forward(self, x, preemph=False, coef=0.97, want_taps=False, layer_taps=False):
                StubEngine.calls += 1
                m = x.float().mean(dim=1)
                s = x.float().std(dim=1)
                return torch.stack([m, s], dim=1)
        model.engine = lambda: StubEngine()
        expect = waves.std(dim=1)
    else:
        from oracle import models_ref as O
        ora = O.build("My_XLSR_AASIST", seed=1024, num_layers=1, order="first")
        model.load_state_dict(ora.state_dict(), strict=True)
        model.rtdf_precision = "fp32"
        with torch.no_grad():
            expect_logits = ora(waves)
            expect = expect_logits[:, 1]
            expect_pe = ora(O.pre_emphasis(waves).reshape(n_utt, n_samples))

    # ---- produce_evaluation_file (main.py:199-221) ------------------------------------------------------------
    tmp = tempfile.mkdtemp()
    save_path = os.path.join(tmp, "scores", "eval.txt")
    ref_main.produce_evaluation_file(FakeSet(), model, device, save_path, batch)
    lines = open(save_path).read().split("\n")
    rows = [ln.split(" ") for ln in lines if ln]
    assert [r[0] for r in rows] == [f"utt_{i:03d}" for i in range(n_utt)], rows
    got = torch.tensor([float(r[1]) for r in rows])
    res["score_file_max_diff"] = float((got - expect).abs().max())
    assert res["score_file_max_diff"] <= (1e-6 if args.mode == "stub" else 1e-4), res

    # ---- Trainer._test (trainer.py:85-132) -----------------------------------------------------------------------
    exp_config = _Cfg(is_pre_emphasis=True, pre_emphasis=0.97)
    tr = ref_trainer.Trainer.__new__(ref_trainer.Trainer)      # __init__ builds the (out-of-scope) augmentation stack
    tr.model = model
    tr.device = device
    tr.preprocessor = prep.PreEmphasis(device, None, exp_config) if args.mode == "cuda" else (lambda x: x)
    weight = torch.tensor([0.9, 0.1])
    tr.loss_fn = torch.nn.CrossEntropyLoss(weight=weight.to(device))
    tr.logger = _Cfg(wandbLog=lambda d: res.setdefault("wandb", {k: float(v) for k, v in d.items()}))
    loader = torch.utils.data.DataLoader(FakeSet(), batch_size=batch, shuffle=False, drop_last=False)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        loss, acc = tr._test(loader)
    if args.mode == "stub":
        logits = torch.stack([waves.mean(dim=1), waves.std(dim=1)], dim=1)
    else:
        logits = expect_pe
    want_loss = 0.0
    for lo in range(0, n_utt, batch):               # trainer.py:108-113: per-batch weighted mean times batch size
        sl = slice(lo, min(lo + batch, n_utt))
        want_loss += float(torch.nn.functional.cross_entropy(logits[sl], labels[sl], weight=weight)) * (sl.stop - sl.start)
    want_loss /= n_utt
    want_acc = 100.0 * float((logits.argmax(1) == labels).float().mean())
    res.update(loss=float(loss), acc=float(acc), want_loss=want_loss, want_acc=want_acc)
    assert abs(loss - want_loss) <= (1e-6 if args.mode == "stub" else 1e-4), res
    assert abs(acc - want_acc) <= 1e-9, res
    if args.mode == "cuda":
        from rtdf_runtime import native
        res["native_launches"] = int(native.load().rtdf_launch_count())
        assert res["native_launches"] > 0, res
    json.dump(res, open(args.out, "w"))


if __name__ == "__main__":
    main()
