"""Drop-in boundary on the GPU: the reference's own scoring callers (main.py:199-221, trainer.py:85-132), unmodified,
drive a 1-layer student through librtdf.so; the score file and loss / accuracy are compared with the CPU oracle."""
import pytest

from tests.test_dropin_host import run_dropin

pytestmark = pytest.mark.gpu


def test_reference_callers_on_cuda(tmp_path):
    res = run_dropin("cuda", tmp_path)
    assert res["native_launches"] > 0
    assert res["score_file_max_diff"] <= 1e-4          # fp32 mode tolerance (north_star)
    assert abs(res["loss"] - res["want_loss"]) <= 1e-4
