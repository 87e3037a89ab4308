"""GPU block / end-to-end parity against the CPU oracle and the committed golden vectors
(produced by the reference's own model files; oracle/check_against_reference.py)."""
import pytest

from tests import e2e_checks as ec

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_backend_block(precision):
    ec.check_backend_block(precision=precision)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_frontend_block(precision):
    ec.check_frontend_block(precision=precision)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_e2e_xlsr_aasist_1s(precision):
    ec.check_e2e(kind="XLSR_AASIST", precision=precision, B=2, N=16000)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_e2e_conformer_1s(precision):
    ec.check_e2e(kind="ConformerModel", precision=precision, B=2, N=16000)


def test_e2e_student_conformer_fixed_call():
    ec.check_e2e(kind="MyModel", precision="bf16", B=2, N=16000, num_layers=2, fixed_call=True)


@pytest.mark.parametrize("name,precision", [
    ("xlsr_aasist_n64000_b2", "bf16"),
    ("xlsr_aasist_n64600_b1", "bf16"),
    ("xlsr_aasist_n16000_b2", "fp32"),
    ("student6_aasist_n64000_b2", "bf16"),
    ("student_mid4_aasist_n16000_b2", "fp32"),
    ("conformer_n64600_b1", "bf16"),
    ("conformer_n16000_b2", "fp32"),
    ("student2_conformer_n16000_b2", "bf16"),
    # fairseq extractor_mode="default": GroupNorm over time after conv-0, plain conv + GELU after (no conv bias)
    ("student2_aasist_groupnorm_n16000_b2", "fp32"),
    ("student2_aasist_groupnorm_n16000_b2", "bf16"),
    ("student2_aasist_groupnorm_n64600_b1", "bf16"),
])
def test_golden_vectors_from_reference(name, precision):
    ec.check_golden(name, precision=precision)


def test_timed_configuration_parity_and_batch_invariance():
    ec.check_timed_configuration()


def test_layer_stack_kernel_streaming_chunks():
    ec.check_layer_stack()


def test_folded_layernorm_mode(monkeypatch):
    """RTDF_LN_FOLD=1 (opt-in): no LayerNorm kernels in the transformer layers -- the projections apply the normalisation
    in their epilogues from per-row statistics the residual GEMMs emit.  Same tolerances as the default path, in the
    streaming-chunk regime (K-split partials + cast kernel), on the wide tiles and through the KD layer taps."""
    import torch
    from tests.util import build_pair
    monkeypatch.setenv("RTDF_LN_FOLD", "1")
    ec.check_e2e(kind="My_XLSR_AASIST", precision="bf16", B=2, N=16000, num_layers=3, order="first")     # 98 rows: split-K
    ec.check_e2e(kind="My_XLSR_AASIST", precision="bf16", B=4, N=64000, num_layers=3, order="first")     # 796 rows: 256-wide tiles
    ora, prod = build_pair("My_XLSR_AASIST", "bf16", num_layers=2, order="first")
    x = ec._waves(2, 16000, seed=3)
    _, taps = prod.engine().forward(x.cuda(), want_taps=True, layer_taps=True)
    with torch.no_grad():
        ref = ora.ssl_model.extract_feat(x)
    assert float((taps["feats"].cpu() - ref).abs().max()) <= ec.FEATS_TOL["bf16"]
    assert taps["layers"].shape == (3, 2, 49, 1024)


def test_ragged_batches_preemphasis_determinism():
    ec.check_ragged_and_quirks(precision="fp32")


def test_utterances_longer_than_one_attention_tile():
    """256 < T <= 512 frames (5.1 .. 10.2 s): the transformer's attention stays on tcgen05 (single 512-column TMEM buffer);
    the AASIST back-end reaches 512 frames (10.2 s; 386 in fp32 mode), the Conformer 512 as well (beyond that the call raises, see
    test_errors_are_loud)."""
    ec.check_e2e("My_XLSR_AASIST", "bf16", B=1, N=88000, num_layers=2, order="first")       # 5.5 s, T = 274
    ec.check_e2e("My_XLSR_AASIST", "bf16", B=2, N=119000, num_layers=2, order="first")      # 7.4 s, T = 371, T' = 123 nodes
    ec.check_e2e("My_XLSR_AASIST", "fp32", B=1, N=119000, num_layers=1, order="first")
    ec.check_e2e("My_XLSR_AASIST", "bf16", B=2, N=164000, num_layers=2, order="first")      # 10.2 s, T = 512, T' = 170 nodes:
    # the transformer's single-buffer attention and the two-box conv slabs of the AASIST encoder at their limit
    ec.check_e2e("MyModel", "bf16", B=1, N=100000, num_layers=2, fixed_call=True)            # 6.25 s, T = 312
    ec.check_e2e("MyModel", "bf16", B=2, N=164000, num_layers=2, fixed_call=True)            # 10.2 s, T = 512 (+ class token)


def test_errors_are_loud():
    import torch
    from tests.util import build_pair
    _, prod = build_pair("My_XLSR_AASIST", "bf16", num_layers=1)
    with pytest.raises(RuntimeError):
        prod(torch.zeros(1, 16000))                       # CPU tensor: no CPU path
    with pytest.raises(RuntimeError):
        prod.train()(torch.zeros(1, 16000, device="cuda"))  # training mode is not accelerated
    prod.eval()
    with pytest.raises((RuntimeError, ValueError)):
        prod(torch.zeros(1, 100, device="cuda"))          # shorter than the conv receptive field
    with pytest.raises(RuntimeError):
        prod(torch.zeros(1, 16000 * 11, device="cuda"))   # T = 549 frames: beyond the tcgen05 attention / AASIST reach of 512
