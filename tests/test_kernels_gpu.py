"""GPU parity tests, kernel by kernel, through the C-ABI (librtdf.so) against fp32 PyTorch references."""
import pytest

from tests import kernel_checks as kc

pytestmark = pytest.mark.gpu


def test_native_library_loaded():
    from tests.util import native
    lib = native().load()
    assert lib.rtdf_launch_count() >= 0


def test_preemph():
    kc.check_preemph()


def test_wave_layernorm():
    kc.check_wave_layernorm()


def test_layernorm_rows():
    kc.check_layernorm()


def test_conv0_ln_gelu():
    kc.check_conv0()


def test_gemm_f32():
    kc.check_gemm_f32()


@pytest.mark.parametrize("variant", [64, 128, 256, 2256])
def test_gemm_bf16_tcgen05(variant):
    kc.check_gemm_bf16(variants=(variant,))


def test_gemm_split_k_partials_and_accumulating_layernorm():
    kc.check_gemm_splitk()


@pytest.mark.parametrize("variant", [256, 2256, 64])
def test_gemm_with_fused_row_layernorm(variant):
    kc.check_gemm_rowln(variants=(variant,))


def test_gelu_epilogue_accuracy():
    kc.check_gelu_epilogue()


@pytest.mark.parametrize("variant", [512, 513, 514, 515, 516])
def test_conv1d_implicit_gemm_ln_gelu(variant):
    kc.check_conv1d_tc(variants=(variant,))


@pytest.mark.parametrize("nsplit", [3, 1])
def test_conv_planes_tcgen05(nsplit):
    kc.check_conv_planes_tc(nsplit=nsplit)


def test_posconv():
    kc.check_posconv()


@pytest.mark.parametrize("impl", [0, 1, 2])
def test_attention(impl):
    kc.check_attention(impls=(impl,))


def test_graph_pool_bit_exact_topk():
    kc.check_graph_pool()


@pytest.mark.parametrize("impl", [0, 1])
def test_graph_attention_rows(impl):
    kc.check_gat_rows(impl)
