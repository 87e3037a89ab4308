"""Drop-in boundary, host side: the reference's own ``main.produce_evaluation_file`` (main.py:199-221) and
``Trainer._test`` (trainer.py:85-132) import and run unmodified with this package's directory ahead of the
reference root on sys.path.  The engine is stubbed (no GPU here); the CUDA version of
the same run is tests/test_dropin_gpu.py."""
import json
import os
import subprocess
import sys

import pytest

from tests.util import ROOT


def run_dropin(mode, tmp_path):
    out = os.path.join(str(tmp_path), "dropin.json")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_runner.py"), "--mode", mode, "--out", out],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.load(open(out))
    if "skipped" in res:
        pytest.skip(res["skipped"])
    return res


def test_reference_callers_run_unmodified_on_the_mirror_package(tmp_path):
    res = run_dropin("stub", tmp_path)
    assert res["score_file_max_diff"] <= 1e-6
    assert abs(res["loss"] - res["want_loss"]) <= 1e-6 and res["acc"] == pytest.approx(res["want_acc"])


def test_package_does_not_shadow_reference_modules():
    """The package directory must not define top-level names the reference owns (utils, config, logger, trainer,
    main, ddp_util) nor make ``data`` a regular package (it has to stay a namespace spanning both roots)."""
    pkg_dir = os.path.join(ROOT, "real-time-deepfake-speech-detection_b200")
    for name in ("utils", "config", "logger", "trainer", "main", "main_kd", "ddp_util"):
        assert not os.path.exists(os.path.join(pkg_dir, name + ".py")), name
        assert not os.path.isdir(os.path.join(pkg_dir, name)), name
    assert not os.path.exists(os.path.join(pkg_dir, "data", "__init__.py"))
    assert sorted(f for f in os.listdir(os.path.join(pkg_dir, "data")) if f.endswith(".py")) == ["preprocess.py"]
