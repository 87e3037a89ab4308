"""CPU: the oracle reproduces the golden vectors that the reference's own files produced
(oracle/check_against_reference.py wrote them), and its building blocks behave as documented."""
import os

import numpy as np
import pytest
import torch

from tests.util import ROOT

GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("name", ["xlsr_aasist_n16000_b2", "student_mid4_aasist_n16000_b2", "conformer_n16000_b2",
                                  "student2_aasist_groupnorm_n16000_b2"])
def test_oracle_matches_reference_golden(name):
    from oracle import models_ref as O
    g = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=True)
    torch.set_num_threads(os.cpu_count() or 1)
    model = O.build(str(g["kind"]), seed=int(g["seed"]), **eval(str(g["kwargs"])))
    x = O.synth_waveforms(int(g["B"]), int(g["N"]), seed=int(g["wave_seed"]))
    taps = {}
    with torch.no_grad():
        y = model(x, taps)
    assert float((y - torch.from_numpy(g["logits"])).abs().max()) <= 2e-5
    assert float((taps["feats"][:, :4, :16] - torch.from_numpy(g["feats_head"])).abs().max()) <= 1e-4
    if "idx_S" in g.files:
        assert torch.equal(taps["idx_S"], torch.from_numpy(g["idx_S"]))
        assert torch.equal(taps["idx_T"], torch.from_numpy(g["idx_T"]))


def test_preemphasis_golden():
    from oracle import models_ref as O
    g = np.load(os.path.join(GOLDEN, "preemph_b3_n4000.npz"))
    y = O.pre_emphasis(O.synth_waveforms(3, 4000, seed=5))
    assert float((y[:, :64] - torch.from_numpy(g["out_head"])).abs().max()) == 0.0
    assert abs(float(y.double().sum()) - float(g["out_sum"])) < 1e-9
    # reflect padding: y[0] = x[0] - 0.97 * x[1]; batch of 1 is squeezed (reference preprocess.py:27)
    x = torch.tensor([[1.0, 2.0, 4.0]])
    assert torch.allclose(O.pre_emphasis(x), torch.tensor([1 - 0.97 * 2, 2 - 0.97 * 1, 4 - 0.97 * 2]))


def test_graph_pool_order_and_size():
    from oracle.aasist_ref import GraphPool
    gp = GraphPool(0.5, 8).eval()
    h = torch.randn(2, 7, 8)
    with torch.no_grad():
        out, idx = gp(h, return_idx=True)
        s = torch.sigmoid(gp.proj(h)).squeeze(-1)
    assert out.shape == (2, 3, 8)                      # floor(7 * 0.5)
    picked = torch.gather(s, 1, idx)
    assert bool((picked[:, :-1] >= picked[:, 1:]).all())  # descending score order
    assert GraphPool(0.5, 8)(torch.randn(1, 1, 8)).shape == (1, 1, 8)  # never fewer than one node


def test_conv_lengths():
    from oracle.wav2vec2_ref import conv_out_lengths
    assert conv_out_lengths(64000) == [12799, 6399, 3199, 1599, 799, 399, 199]
    assert conv_out_lengths(64600)[-1] == 201 and conv_out_lengths(16000)[-1] == 49


def test_student_conformer_typeerror_is_reproduced():
    from oracle import models_ref as O
    m = O.build("MyModel", num_layers=1)
    with pytest.raises(TypeError):
        m(torch.zeros(1, 16000))
