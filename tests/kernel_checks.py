"""Per-kernel parity checks of the CUDA path (through the C-ABI) against plain fp32 PyTorch
references of the same op.  Each check returns a dict of measured errors and raises
AssertionError with the numbers when a tolerance (written next to each assert) is exceeded.
Used by tests/test_kernels_gpu.py and tools/gpu_diag.py."""
import math

import torch
import torch.nn.functional as F

from tests.util import P, call, gemm_bf16, gemm_f32, stream

DEV = "cuda"
_KEEP = []


def dev(t):
    """Copy to the device and keep the tensor alive (a bare P(dev(x)) would free it before the launch)."""
    d = t.to(DEV)
    _KEEP.append(d)
    if len(_KEEP) > 256:
        torch.cuda.synchronize()
        del _KEEP[:128]
    return d



class Guarded:
    """Output buffer with canary margins (compute-sanitizer is closed on this GPU pool, so out-of-bounds WRITES are caught
    here): the kernel gets a view into the middle of a larger allocation; check() asserts the margins are untouched."""
    PAD = 4096          # elements on either side
    CANARY = -7777.0

    def __init__(self, shape, dtype=torch.float32, fill=0.0):
        n = 1
        for d in shape:
            n *= d
        self.raw = torch.full((n + 2 * self.PAD,), self.CANARY, dtype=dtype, device=DEV)
        self.t = self.raw[self.PAD:self.PAD + n].view(*shape)
        self.t.fill_(fill)

    def check(self, what=""):
        lo, hi = self.raw[:self.PAD], self.raw[-self.PAD:]
        ok = bool((lo == self.CANARY).all()) and bool((hi == self.CANARY).all())
        assert ok, f"out-of-bounds write next to the output of {what}"
        return self.t


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12))


def check_preemph():
    from oracle.models_ref import pre_emphasis, synth_waveforms
    out = {}
    for B, N in ((3, 4000), (2, 64600), (1, 4001)):
        x = synth_waveforms(B, N, seed=5)
        ref = pre_emphasis(x).reshape(B, N)
        xd = x.to(DEV)
        y = torch.empty_like(xd)
        call("rtdf_preemph", P(xd), P(y), B, N, 0.97, stream())
        d = float((y.cpu() - ref).abs().max())
        out[f"{B}x{N}"] = d
        assert d <= 1e-6, out      # fp32, FMA contraction only
    return out


def check_wave_layernorm():
    x = torch.randn(3, 16000) * 0.3 + 0.1
    xd = x.to(DEV)
    y = torch.empty_like(xd)
    call("rtdf_wave_layernorm", P(xd), P(y), 3, 16000, 1e-5, stream())
    ref = F.layer_norm(x, (16000,))
    d = float((y.cpu() - ref).abs().max())
    assert d <= 2e-5, d
    return {"max_abs": d}


def check_layernorm():
    out = {}
    g = torch.Generator().manual_seed(0)
    for rows, C in ((37, 1024), (130, 512), (201, 144), (9, 128)):
        x = torch.randn(rows, C, generator=g) * 2 + 0.5
        gamma = 1 + 0.1 * torch.randn(C, generator=g)
        beta = 0.1 * torch.randn(C, generator=g)
        ref = F.layer_norm(x, (C,), gamma, beta, 1e-5)
        for in_bf16 in (0, 1):
            xin = x.to(DEV).to(torch.bfloat16) if in_bf16 else x.to(DEV)
            r = F.layer_norm(xin.float().cpu(), (C,), gamma, beta, 1e-5) if in_bf16 else ref
            o32 = torch.empty(rows, C, dtype=torch.float32, device=DEV)
            o16 = torch.empty(rows, C, dtype=torch.bfloat16, device=DEV)
            call("rtdf_layernorm_rows", P(xin), in_bf16, rows, C, P(dev(gamma)), P(dev(beta)), 1e-5, 0, P(o32),
                 P(o16), stream())
            d32 = float((o32.cpu() - r).abs().max())
            d16 = float((o16.float().cpu() - r).abs().max())
            out[f"{rows}x{C}_bf16in{in_bf16}"] = (d32, d16)
            assert d32 <= 2e-5, out      # fp32 statistics
            assert d16 <= 0.04, out      # bf16 output rounding of O(4) values
        # fused GELU
        o32 = torch.empty(rows, C, dtype=torch.float32, device=DEV)
        call("rtdf_layernorm_rows", P(dev(x)), 0, rows, C, P(dev(gamma)), P(dev(beta)), 1e-5, 1, P(o32), None, stream())
        d = float((o32.cpu() - F.gelu(ref)).abs().max())
        assert d <= 2e-5, d
    return out


def check_conv0():
    g = torch.Generator().manual_seed(1)
    out = {}
    for B, N in ((2, 16000), (1, 4003)):
        wav = torch.randn(B, N, generator=g) * 0.1
        w = torch.randn(512, 1, 10, generator=g) * 0.4
        b = torch.randn(512, generator=g) * 0.1
        gamma = 1 + 0.1 * torch.randn(512, generator=g)
        beta = 0.1 * torch.randn(512, generator=g)
        y = F.conv1d(wav.unsqueeze(1), w, b, stride=5)            # (B,512,L)
        ref = F.gelu(F.layer_norm(y.transpose(1, 2), (512,), gamma, beta, 1e-5))   # (B,L,512)
        L = ref.shape[1]
        wt = w[:, 0, :].t().contiguous().to(DEV)
        o32 = torch.empty(B, L, 512, dtype=torch.float32, device=DEV)
        call("rtdf_conv0_ln_gelu", P(dev(wav)), B, N, P(wt), P(dev(b)), P(dev(gamma)), P(dev(beta)), 1e-5,
             P(o32), None, stream())
        o16 = torch.empty(B, L, 512, dtype=torch.bfloat16, device=DEV)
        call("rtdf_conv0_ln_gelu", P(dev(wav)), B, N, P(wt), P(dev(b)), P(dev(gamma)), P(dev(beta)), 1e-5,
             None, P(o16), stream())
        d32 = float((o32.cpu() - ref).abs().max())
        d16 = float((o16.float().cpu() - ref).abs().max())
        out[f"{B}x{N}"] = (d32, d16)
        assert d32 <= 5e-5, out
        assert d16 <= 0.03, out
    return out


def check_conv0_tc(shapes=((2, 16000), (1, 4003), (3, 64600))):
    """conv-0 as a tcgen05 implicit GEMM (hi/lo-split K = 32 operands, LN + GELU epilogue) against the fp32 reference.
    The split keeps the conv itself at ~2^-16 relative, so the error is the bf16 rounding of the output."""
    from tests.util import native
    lib = native().load()
    g = torch.Generator().manual_seed(41)
    out = {}
    for B, N in shapes:
        wav = torch.randn(B, N, generator=g) * 0.1 + 0.03
        w = torch.randn(512, 1, 10, generator=g) * 0.4
        b = torch.randn(512, generator=g) * 0.1
        gamma = 1 + 0.1 * torch.randn(512, generator=g)
        beta = 0.1 * torch.randn(512, generator=g)
        if B * N > 2_000_000:
            _no_tf32()
            y = F.conv1d(wav.to(DEV).unsqueeze(1), w.to(DEV), b.to(DEV), stride=5)
            ref = F.gelu(F.layer_norm(y.transpose(1, 2), (512,), gamma.to(DEV), beta.to(DEV), 1e-5)).cpu()
        else:
            y = F.conv1d(wav.unsqueeze(1), w, b, stride=5)
            ref = F.gelu(F.layer_norm(y.transpose(1, 2), (512,), gamma, beta, 1e-5))
        L = ref.shape[1]
        scratch = torch.empty(int(lib.rtdf_conv0_tc_scratch_bytes(B, N)), dtype=torch.uint8, device=DEV)
        go = Guarded((B, L, 512), torch.bfloat16)
        o16 = go.t
        call("rtdf_conv0_tc_ln_gelu", P(dev(wav)), B, N, P(dev(w.reshape(512, 10).contiguous())), P(dev(b)), P(dev(gamma)),
             P(dev(beta)), 1e-5, P(scratch), P(o16), stream())
        go.check(f"rtdf_conv0_tc_ln_gelu {B}x{N}")
        got = o16.float().cpu()
        d16 = float((got - ref).abs().max())
        # the SIMT kernel's bf16 output of the same layer: both round the same fp32 value, so they agree to one bf16 ulp
        o_simt = torch.zeros(B, L, 512, dtype=torch.bfloat16, device=DEV)
        call("rtdf_conv0_ln_gelu", P(dev(wav)), B, N, P(dev(w[:, 0, :].t().contiguous())), P(dev(b)), P(dev(gamma)),
             P(dev(beta)), 1e-5, None, P(o_simt), stream())
        d_simt = float((got - o_simt.float().cpu()).abs().max())
        out[f"{B}x{N}"] = (d16, d_simt)
        assert d16 <= 0.03, out          # bf16 rounding of O(4) values
        assert d_simt <= 0.04, out
    return out


def check_conv0_groupnorm():
    """conv-0 of the group-norm feature encoder (fairseq extractor_mode="default"): conv -> GroupNorm(512, 512) over
    time -> GELU, with and without a conv bias, ragged chunk boundaries (L1 = 3199, 799, 256, 257)."""
    from tests.util import native
    lib = native().load()
    g = torch.Generator().manual_seed(31)
    out = {}
    for B, N, with_bias in ((2, 16000, False), (3, 4003, True), (1, 1285, False), (2, 1290, True)):
        wav = torch.randn(B, N, generator=g) * 0.1 + 0.02
        w = torch.randn(512, 1, 10, generator=g) * 0.4
        b = torch.randn(512, generator=g) * 0.1 if with_bias else None
        gamma = 1 + 0.1 * torch.randn(512, generator=g)
        beta = 0.1 * torch.randn(512, generator=g)
        y = F.conv1d(wav.unsqueeze(1), w, b, stride=5)                              # (B,512,L)
        ref = F.gelu(F.group_norm(y, 512, gamma, beta, 1e-5)).transpose(1, 2)       # (B,L,512)
        L = ref.shape[1]
        ws = torch.empty(int(lib.rtdf_conv0_gn_workspace_floats(B, N)), device=DEV)
        wt = w[:, 0, :].t().contiguous().to(DEV)
        o32 = torch.full((B, L, 512), float("nan"), device=DEV)
        o16 = torch.zeros(B, L, 512, dtype=torch.bfloat16, device=DEV)
        bd = dev(b) if with_bias else None
        call("rtdf_conv0_gn_gelu", P(dev(wav)), B, N, P(wt), P(bd), P(dev(gamma)), P(dev(beta)), 1e-5, P(ws), P(o32), None, stream())
        call("rtdf_conv0_gn_gelu", P(dev(wav)), B, N, P(wt), P(bd), P(dev(gamma)), P(dev(beta)), 1e-5, P(ws), None, P(o16), stream())
        d32 = float((o32.cpu() - ref).abs().max())
        d16 = float((o16.float().cpu() - ref).abs().max())
        out[f"{B}x{N}_bias{int(with_bias)}"] = (d32, d16)
        assert d32 <= 5e-5, out
        assert d16 <= 0.03, out
    return out


def check_gemm_f32():
    g = torch.Generator().manual_seed(2)
    out = {}
    for M, N, K in ((70, 130, 96), (257, 64, 1024), (33, 2, 144)):
        A = torch.randn(M, K, generator=g)
        W = torch.randn(N, K, generator=g) / math.sqrt(K)
        b = torch.randn(N, generator=g)
        R = torch.randn(M, N, generator=g)
        ref = F.gelu(A @ W.t() + b) * 0.5 + R
        o = gemm_f32(A.to(DEV), W.to(DEV), b.to(DEV), act=1, scale=0.5, resid=R.to(DEV))
        d = float((o.cpu() - ref).abs().max())
        out[f"{M}x{N}x{K}"] = d
        assert d <= 2e-5, out
    return out


def check_gemm_bf16(variants=(64, 128, 256)):
    g = torch.Generator().manual_seed(3)
    out = {}
    shapes = ((128, 256, 64), (300, 256, 128), (1000, 3072, 1024), (200, 144, 144), (131, 576, 144), (77, 1024, 4096),
              (402, 128, 1024), (5000, 1024, 256))     # the last one: more tiles than SMs (persistent loop, both stages)
    for M, N, K in shapes:
        A = (torch.randn(M, K, generator=g)).to(torch.bfloat16)
        W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16)
        b = torch.randn(N, generator=g)
        R = torch.randn(M, N, generator=g)
        base = A.float() @ W.float().t()
        for v in variants:
            o = gemm_bf16(A.to(DEV), W.to(DEV), variant=v)
            d0 = float((o.cpu() - base).abs().max())
            o32, o16 = gemm_bf16(A.to(DEV), W.to(DEV), b.to(DEV), act=1, scale=0.5, resid=R.to(DEV), out="both", variant=v)
            ref = F.gelu(base + b) * 0.5 + R
            d1 = float((o32.cpu() - ref).abs().max())
            d2 = float((o16.float().cpu() - ref).abs().max())
            o_in = gemm_bf16(A.to(DEV), W.to(DEV), b.to(DEV), act=1, scale=0.5, resid=R.to(DEV), variant=v, inplace=True)
            d3 = float((o_in.cpu() - ref).abs().max())
            o_b = gemm_bf16(A.to(DEV), W.to(DEV), b.to(DEV), act=1, scale=0.5, out="bf16", variant=v)
            d4 = float((o_b.float().cpu() - F.gelu(base + b) * 0.5).abs().max())
            out[f"{M}x{N}x{K}_v{v}"] = (d0, d1, d2, d3, d4)
            assert d0 <= 2e-3, out     # fp32 accumulation of exact bf16 products, different summation order
            assert d1 <= 2e-3, out
            assert d2 <= 0.05, out     # bf16 output rounding
            assert d3 <= 2e-3, out     # in-place residual update (TMA reduce-add)
            assert d4 <= 0.05, out     # bf16-only output (TMA store)
    return out


def check_gemm_splitk():
    """Skinny residual GEMM + LayerNorm of the streaming-chunk regime: K split over idle SMs, partial sums folded into
    the residual stream in split order by the LayerNorm that follows (bit-reproducible)."""
    from tests.util import native
    lib = native().load()
    g = torch.Generator().manual_seed(13)
    out = {}
    assert lib.rtdf_gemm_plan_splits(12736, 1024, 4096) == 1          # large batch: never split
    for M, K in ((49, 4096), (49, 1024), (199, 4096), (392, 1024), (8, 1024), (500, 4096)):
        N = 1024
        S = lib.rtdf_gemm_plan_splits(M, N, K)
        assert 2 <= S <= 8, (M, K, S)
        A = (torch.randn(M, K, generator=g)).to(torch.bfloat16)
        W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16)
        b = torch.randn(N, generator=g)
        x0 = torch.randn(M, N, generator=g) * 2
        gamma = 1 + 0.1 * torch.randn(N, generator=g)
        beta = 0.1 * torch.randn(N, generator=g)
        xr = x0 + A.float() @ W.float().t() + b
        ref = F.layer_norm(xr, (N,), gamma, beta, 1e-5)
        runs = []
        for rep in range(3):
            parts = torch.full((S, M, N), float("nan"), device=DEV)
            x = x0.to(DEV)
            o16 = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
            call("rtdf_gemm_bf16_splitk", P(dev(A)), P(dev(W)), M, N, K, P(dev(b)), P(parts), stream())
            call("rtdf_layernorm_accum_rows", P(x), P(parts), S, M, P(dev(gamma)), P(dev(beta)), 1e-5, None, P(o16), stream())
            x2 = x0.to(DEV)
            o32 = torch.empty(M, N, device=DEV)
            call("rtdf_layernorm_accum_rows", P(x2), P(parts), S, M, P(dev(gamma)), P(dev(beta)), 1e-5, P(o32), None, stream())
            runs.append((x.cpu(), o16.cpu(), o32.cpu()))
        dx = float((runs[0][0] - xr).abs().max())
        d32 = float((runs[0][2] - ref).abs().max())
        d16 = float((runs[0][1].float() - ref).abs().max())
        out[f"{M}x{N}x{K}_S{S}"] = (dx, d32, d16)
        assert dx <= 2e-3 and d32 <= 2e-3 and d16 <= 0.04, out
        for r in runs[1:]:           # bit-reproducible: no atomics, fixed summation order
            assert torch.equal(r[0], runs[0][0]) and torch.equal(r[1], runs[0][1]) and torch.equal(r[2], runs[0][2])
    return out


def check_gemm_rowln(variants=(256, 2256)):
    """x += A W^T + b followed by the fused per-row-block LayerNorm (out_proj -> LN / fc2 -> LN pattern)."""
    g = torch.Generator().manual_seed(13)
    out = {}
    for M, N, K in ((12736, 1024, 1024), (300, 1024, 256), (49, 1024, 4096), (1000, 512, 128), (5000, 1024, 64)):
        A = torch.randn(M, K, generator=g).to(torch.bfloat16)
        W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16)
        b = torch.randn(N, generator=g)
        x0 = torch.randn(M, N, generator=g) * 2
        gamma = 1 + 0.1 * torch.randn(N, generator=g)
        beta = 0.1 * torch.randn(N, generator=g)
        xr = x0 + A.float() @ W.float().t() + b
        ref = F.layer_norm(xr, (N,), gamma, beta, 1e-5)
        Ad, Wd, bd, gd, btd = dev(A), dev(W), dev(b), dev(gamma), dev(beta)
        for v in variants:
            cnt = torch.zeros((M + 127) // 128, dtype=torch.int32, device=DEV)
            for rep in range(3):          # repeated launches: counters must come back to zero, no race on the row blocks
                x = x0.to(DEV)
                o16 = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
                o32 = torch.full((M, N), float("nan"), dtype=torch.float32, device=DEV)
                call("rtdf_gemm_bf16_rowln", P(Ad), P(Wd), M, N, K, P(bd), P(x), P(gd), P(btd), 1e-5, P(o16), P(o32), P(cnt),
                     v, stream())
                dx = float((x.cpu() - xr).abs().max())
                d32 = float((o32.cpu() - ref).abs().max())
                d16 = float((o16.float().cpu() - ref).abs().max())
                out[f"{M}x{N}x{K}_v{v}_rep{rep}"] = (dx, d32, d16)
                assert int(cnt.abs().sum()) == 0, out
                assert dx <= 2e-3 and d32 <= 2e-3, out        # fp32 accumulation order only
                assert d16 <= 0.04, out                       # bf16 rounding of O(4) values
    return out


def check_gemm_lnfold(variants=(256, 2256)):
    """LayerNorm folded into the GEMMs around it (how the bf16 transformer layers run): the residual GEMM emits x, bf16(x)
    and per-row partial statistics; the next projection evaluates LN(x) W^T + b as rstd (bf16(x) W'^T - mean c) + d."""
    g = torch.Generator().manual_seed(17)
    out = {}
    N = 1024
    for M, K, N2 in ((12736, 1024, 3072), (300, 4096, 4096), (49, 1024, 1024), (2500, 256, 512)):
        A = torch.randn(M, K, generator=g).to(torch.bfloat16)
        W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16)
        b = torch.randn(N, generator=g)
        x0 = torch.randn(M, N, generator=g) * 2 + 0.3
        gamma = 1 + 0.1 * torch.randn(N, generator=g)
        beta = 0.1 * torch.randn(N, generator=g)
        W2 = torch.randn(N2, N, generator=g) / math.sqrt(N)
        b2 = torch.randn(N2, generator=g)
        xr = x0 + A.float() @ W.float().t() + b
        ref = F.gelu(F.layer_norm(xr, (N,), gamma, beta, 1e-5) @ W2.t() + b2)
        wf = torch.empty(N2, N, dtype=torch.bfloat16, device=DEV)
        cv, dv = torch.empty(N2, device=DEV), torch.empty(N2, device=DEV)
        call("rtdf_fold_ln_weight", P(dev(W2)), P(dev(gamma)), P(dev(beta)), P(dev(b2)), N2, N, P(wf), P(cv), P(dv), stream())
        want_wf = (W2 * gamma).to(torch.bfloat16)
        assert torch.equal(wf.cpu(), want_wf), "folded weight"
        assert float((cv.cpu() - want_wf.float().sum(1)).abs().max()) <= 1e-3
        assert float((dv.cpu() - (b2 + W2 @ beta)).abs().max()) <= 1e-4
        Ad, Wd, bd = dev(A), dev(W), dev(b)
        for v in variants:
            runs = []
            for rep in range(2):
                x = x0.to(DEV)
                xb = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
                stats = torch.full((M, 8, 2), float("nan"), device=DEV)
                call("rtdf_gemm_bf16_xres", P(Ad), P(Wd), M, N, K, P(bd), P(x), P(xb), P(stats), v, stream())
                o16 = torch.zeros(M, N2, dtype=torch.bfloat16, device=DEV)
                call("rtdf_gemm_bf16_lnfold", P(xb), P(wf), M, N2, N, P(cv), P(dv), P(stats), 1e-5, 1, None, P(o16),
                     v, stream())
                runs.append((x.cpu(), xb.cpu(), stats.cpu(), o16.cpu()))
            x, xb, stats, o16 = runs[0]
            dx = float((x - xr).abs().max())
            assert dx <= 2e-3, (M, K, v, dx)                        # fp32 accumulation order only
            assert torch.equal(xb, x.to(torch.bfloat16)), "bf16 copy of the residual stream"
            ssum, ssq = stats[:, :, 0].sum(1), stats[:, :, 1].sum(1)
            assert float((ssum - x.sum(1)).abs().max()) <= 2e-2 and float((ssq / x.pow(2).sum(1) - 1).abs().max()) <= 1e-4
            d16 = float((o16.float() - ref).abs().max())
            out[f"{M}x{K}->{N2}_v{v}"] = (dx, d16)
            assert d16 <= 0.06, out              # bf16 rounding of x, of W' and of the O(4) outputs
            for a_, b_ in zip(runs[0], runs[1]):     # no atomics anywhere: bit-reproducible
                assert torch.equal(a_, b_) or (torch.isnan(a_) == torch.isnan(b_)).all()
        # streaming-chunk producer: K-split partial sums folded by the cast kernel, consumer on 64-wide tiles (direct stores)
        if M <= 512:
            from tests.util import native
            lib = native().load()
            S = lib.rtdf_gemm_plan_splits(M, N, K)
            parts = torch.empty(S, M, N, device=DEV)
            call("rtdf_gemm_bf16_splitk", P(Ad), P(Wd), M, N, K, P(bd), P(parts), stream())
            x = x0.to(DEV)
            xb = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
            stats = torch.full((M, 8, 2), float("nan"), device=DEV)
            call("rtdf_cast_stats_rows", P(x), P(parts), S, M, P(xb), P(stats), stream())
            assert float((x.cpu() - xr).abs().max()) <= 2e-3
            assert torch.equal(xb.cpu(), x.cpu().to(torch.bfloat16))
            assert float(stats[:, 1:, :].abs().max()) == 0.0
            o16 = torch.zeros(M, N2, dtype=torch.bfloat16, device=DEV)
            call("rtdf_gemm_bf16_lnfold", P(xb), P(wf), M, N2, N, P(cv), P(dv), P(stats), 1e-5, 1, None, P(o16), 64, stream())
            d16 = float((o16.float().cpu() - ref).abs().max())
            out[f"{M}x{K}->{N2}_skinny"] = d16
            assert d16 <= 0.06, out
    return out


def check_gelu_epilogue():
    """The tensor-core epilogue GELUs against the exact erf GELU on a dense grid (identity GEMM, fp32 output).
    act 1 = sigmoid-of-quintic fit (product default), 4 = hardware-tanh form, 5 = A&S 7.1.26 erf."""
    M, K = 2048, 64
    xs = torch.linspace(-12.0, 12.0, M * K).to(torch.bfloat16).reshape(M, K)     # bf16-exact inputs
    eye = torch.eye(K).to(torch.bfloat16)
    ref = F.gelu(xs.double()).float()
    out = {}
    for act, tol in ((1, 4e-5), (4, 1.5e-3), (5, 2e-6)):
        o = gemm_bf16(xs.to(DEV), eye.to(DEV), act=act, variant=64)
        d = float((o.cpu() - ref).abs().max())
        out[f"act{act}"] = d
        assert d <= tol, out
    return out


def _split(t):
    hi = t.to(torch.bfloat16)
    return hi, (t - hi.float()).to(torch.bfloat16)


CONV_PLANES_CASES = [  # (kind, ci, co, B, W)
    ("conv1", 32, 32, 2, 16), ("conv1", 32, 64, 1, 66), ("conv1", 64, 64, 2, 66), ("conv2", 32, 32, 2, 66),
    ("conv2", 64, 64, 3, 67), ("ds", 32, 64, 2, 66), ("lin", 64, 128, 2, 66), ("lin", 128, 64, 2, 16),
    ("conv2", 64, 64, 1, 96),
    # planes wider than 124 columns: the 128 + W + 4 row slab arrives as two TMA boxes (utterances of 7.5 .. 10.2 s)
    ("conv1", 64, 64, 1, 170), ("conv2", 32, 32, 2, 150), ("ds", 32, 64, 1, 170), ("conv2", 64, 64, 1, 133),
]
# the timed configuration: B = 64, T' = 66 -> 64 * 44 * 68 = 191,488 plane rows = 1,496 row tiles on 148 persistent CTAs
CONV_PLANES_CASES_TIMED = [("conv1", 64, 64, 64, 66), ("conv1", 32, 64, 64, 66), ("conv2", 64, 64, 64, 66),
                           ("conv2", 32, 32, 64, 66), ("ds", 32, 64, 64, 66), ("lin", 64, 128, 64, 66),
                           ("lin", 128, 64, 64, 66)]


def check_conv_planes_tc(nsplit=3, cases=None):
    """Shifted-row tcgen05 conv on zero-padded channels-last planes against F.conv2d (fp32) for the three
    Residual_block convolutions (aasist_modules.py:340-397) and the 1x1 attention convs (xlsr_aasist.py:103)."""
    import ctypes
    g = torch.Generator().manual_seed(11)
    out = {}
    Hp = 44
    for kind, ci, co, B, W in (cases or CONV_PLANES_CASES):
        Wp = W + 2
        rows = B * Hp * Wp
        Hin = 43 if kind == "conv2" else 42
        x = torch.randn(B, ci, Hin, W, generator=g)
        plane = torch.zeros(B, Hp, Wp, ci)
        r0 = 0 if kind == "conv2" else 1            # 43-row tensors sit at hp = h, 42-row tensors at hp = h + 1
        plane[:, r0:r0 + Hin, 1:W + 1, :] = x.permute(0, 2, 3, 1)
        bias = torch.randn(co, generator=g) * 0.1
        s1 = 1 + 0.1 * torch.randn(co, generator=g)
        t1 = 0.1 * torch.randn(co, generator=g)
        s2 = 1 + 0.1 * torch.randn(co, generator=g)
        t2 = 0.1 * torch.randn(co, generator=g)
        if kind == "lin":
            w = torch.randn(co, ci, generator=g) / math.sqrt(ci)
            ref = F.conv2d(x, w[:, :, None, None], bias)
            nb = (ci + 63) // 64
            wch = torch.stack([w[:, 64 * j:64 * j + min(ci, 64)] for j in range(nb)])           # [chunk][co][kw]
            shift, sub, h0, nrow = [0] * nb, list(range(nb)), 1, 42
            hp_lo, hp_hi = 1, 42
        else:
            KH = 1 if kind == "ds" else 2
            w = torch.randn(co, ci, KH, 3, generator=g) / math.sqrt(ci * KH * 3)
            pad_h = 1 if kind == "conv1" else 0
            ref = F.conv2d(x, w, bias, padding=(pad_h, 1))
            wch = torch.stack([w[:, :, kh, kw] for kh in range(KH) for kw in range(3)])          # [tap][co][ci]
            if kind == "conv1":
                shift = [kh * Wp + kw - 1 for kh in range(2) for kw in range(3)]
                h0, nrow, hp_lo, hp_hi = 0, 43, 0, 42
            elif kind == "conv2":
                shift = [(kh - 1) * Wp + kw - 1 for kh in range(2) for kw in range(3)]
                h0, nrow, hp_lo, hp_hi = 1, 42, 1, 42
            else:
                shift = [kw - 1 for kw in range(3)]
                h0, nrow, hp_lo, hp_hi = 1, 42, 1, 42
            sub = [0] * len(shift)
        resid = torch.zeros(B, Hp, Wp, co)
        resid[:, h0:h0 + nrow, 1:W + 1, :] = torch.randn(B, nrow, W, co, generator=g)
        expect = F.selu(ref * s1[None, :, None, None] + t1[None, :, None, None])
        expect = expect.permute(0, 2, 3, 1) + resid[:, h0:h0 + nrow, 1:W + 1, :]
        expect = F.selu(expect * s2 + t2)
        full = torch.zeros(B, Hp, Wp, co)
        full[:, h0:h0 + nrow, 1:W + 1, :] = expect
        xh, xl = _split(plane.reshape(rows, ci))
        wh, wl = _split(wch.contiguous())
        o32 = torch.full((rows, co), float("nan"), device=DEV)
        oh = torch.empty(rows, co, dtype=torch.bfloat16, device=DEV)
        ol = torch.empty(rows, co, dtype=torch.bfloat16, device=DEV)
        n = len(shift)
        sh = (ctypes.c_int * n)(*shift)
        sb = (ctypes.c_int * n)(*sub)
        call("rtdf_conv_planes_tc", P(dev(xh)), P(dev(xl)), ci, rows, Hp, Wp, P(dev(wh)), P(dev(wl)), co, n, sh, sb,
             hp_lo, hp_hi, P(dev(bias)), P(dev(s1)), P(dev(t1)), 3, P(dev(resid.reshape(rows, co))), P(dev(s2)),
             P(dev(t2)), 3, P(o32), P(oh), P(ol), nsplit, stream())
        got = o32.cpu().reshape(B, Hp, Wp, co)
        d = float((got - full).abs().max())
        d16 = float(((oh.float() + ol.float()).cpu().reshape(B, Hp, Wp, co) - full).abs().max())
        out[f"{kind}_{ci}to{co}_B{B}_W{W}"] = (d, d16)
        assert torch.isfinite(got).all(), out
        assert d <= (2e-4 if nsplit == 3 else 5e-2), out     # 3-term split: ~2^-16 relative per product
        assert d16 <= d + 1e-4, out                           # hi + lo reproduces the fp32 plane to ~2^-17
    return out


def _conv_ref(x, w, b, gamma, beta, stride):
    y = F.conv1d(x.transpose(1, 2), w, b, stride=stride)             # (B,512,L)
    return F.gelu(F.layer_norm(y.transpose(1, 2), (512,), gamma, beta, 1e-5))


CONV1D_SHAPES = ((2, 799, 3), (1, 403, 2), (3, 130, 3), (2, 12799, 3))
# the timed configuration: conv-1 (64 x 12,799 -> 6,399 frames: 3,200 row tiles, 1,600 CTA pairs) and conv-5 (64 x 799 -> 399)
CONV1D_SHAPES_TIMED = ((64, 12799, 3), (64, 799, 2))


def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")


def check_conv1d_tc(variants=(512, 513), shapes=CONV1D_SHAPES):
    g = torch.Generator().manual_seed(4)
    out = {}
    for B, L, k in shapes:
        x = torch.randn(B, L, 512, generator=g).to(torch.bfloat16)
        w = (torch.randn(512, 512, k, generator=g) / math.sqrt(512 * k)).to(torch.bfloat16)
        b = torch.randn(512, generator=g) * 0.1
        gamma = 1 + 0.1 * torch.randn(512, generator=g)
        beta = 0.1 * torch.randn(512, generator=g)
        if B * L > 100000:       # 0.6 TFLOP of fp32 reference: plain PyTorch fp32 ops on the GPU (TF32 off), in slices
            _no_tf32()
            ref = torch.cat([_conv_ref(x[i:i + 8].to(DEV).float(), w.to(DEV).float(), b.to(DEV), gamma.to(DEV),
                                       beta.to(DEV), 2).cpu() for i in range(0, B, 8)])
        else:
            ref = _conv_ref(x.float(), w.float(), b, gamma, beta, 2)
        Lo = ref.shape[1]
        wp = w.permute(0, 2, 1).contiguous().to(DEV)               # [co][k][ci]
        for v in variants:
            gy = Guarded((B, Lo, 512), torch.bfloat16)
            y = gy.t
            call("rtdf_conv1d_ln_gelu_bf16", P(dev(x)), B, L, k, 2, P(wp), P(dev(b)), P(dev(gamma)),
                 P(dev(beta)), 1e-5, P(y), v, stream())
            gy.check(f"rtdf_conv1d_ln_gelu_bf16 variant {v} B{B} L{L}")
            d = float((y.float().cpu() - ref).abs().max())
            out[f"B{B}_L{L}_k{k}_v{v}"] = d
            assert d <= 0.04, out    # bf16 output rounding of O(4) values
    return out


def _posconv_ref(x, w, bias):
    B, T, C = x.shape
    if B * T > 4000:        # 0.2 TFLOP: plain PyTorch fp32 ops on the GPU (TF32 off)
        _no_tf32()
        x, w, bias = x.to(DEV), w.to(DEV), bias.to(DEV)
    y = F.conv1d(x.transpose(1, 2), w, bias, padding=64, groups=16)[:, :, :T]
    return (x + F.gelu(y).transpose(1, 2)).cpu()


POSCONV_SHAPES = ((2, 199), (1, 49), (2, 130), (3, 128), (2, 257), (1, 1), (2, 520))
POSCONV_SHAPES_TIMED = ((64, 199),)      # 1,024 (utterance, group) items on 148 persistent CTAs: ~7 items per CTA


def check_posconv(shapes=POSCONV_SHAPES):
    g = torch.Generator().manual_seed(5)
    out = {}
    for B, T in shapes:
        x = torch.randn(B, T, 1024, generator=g)
        w = torch.randn(1024, 64, 128, generator=g) / math.sqrt(64 * 128)
        bias = torch.randn(1024, generator=g) * 0.1
        wp = w.permute(0, 2, 1).reshape(1024, 8192).contiguous()    # [co][k*64+ci]
        ref32 = _posconv_ref(x, w, bias)
        xo = x.clone().to(DEV)
        call("rtdf_posconv_f32", P(xo), P(dev(x)), B, T, P(dev(wp)), P(dev(bias)), stream())
        d32 = float((xo.cpu() - ref32).abs().max())
        xb = x.to(torch.bfloat16)
        wb = wp.to(torch.bfloat16)
        ref16 = x + (_posconv_ref(xb.float(), wb.float().reshape(1024, 128, 64).permute(0, 2, 1), bias) - xb.float())
        d16 = 0.0
        for impl in (0, 1):      # 0 = slab-resident kernel, 1 = tap-shifted GEMM
            gx = Guarded((B, T, 1024))
            xo2 = gx.t
            xo2.copy_(x)
            call("rtdf_posconv_bf16", P(xo2), P(dev(xb)), B, T, P(dev(wb)), P(dev(bias)), impl, stream())
            gx.check(f"rtdf_posconv_bf16 impl {impl} B{B} T{T}")
            d16 = max(d16, float((xo2.cpu() - ref16).abs().max()))
        out[f"B{B}_T{T}"] = (d32, d16)
        assert d32 <= 5e-5, out
        assert d16 <= 2e-3, out
    return out


def _attn_ref(qkv, B, T, H):
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    p = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B * T, H * 64)


ATTN_SHAPES = ((2, 199, 16), (1, 49, 16), (2, 201, 4), (1, 128, 2), (1, 256, 2), (3, 17, 1))
# the timed configuration (BASELINE configs[2]): 2,048 (utterance, head, query tile) items on <= 148 persistent CTAs, so
# every CTA loops ~14 times over its smem ring / TMEM buffers / mbarrier phases
ATTN_SHAPES_TIMED = ((64, 199, 16), (64, 201, 16))
# 256 < T <= 512 frames (5.1 .. 10.2 s): single 512-column TMEM buffer, K / V in two TMA boxes, S as two MMAs per k-step;
# (40, 300, 16) gives 1,920 items on 148 CTAs (barrier phases of the one-stage ring flip many times)
ATTN_SHAPES_LONG = ((1, 257, 2), (2, 300, 4), (1, 400, 16), (1, 512, 2), (2, 272, 1), (40, 300, 16))


def check_attention(impls=(0, 1), shapes=ATTN_SHAPES):
    g = torch.Generator().manual_seed(6)
    out = {}
    for B, T, H in shapes:
        qkv = torch.randn(B * T, 3 * H * 64, generator=g)
        qkv[:, : H * 64] *= 0.25
        ref32 = _attn_ref(qkv, B, T, H)
        if T <= 400:      # the fp32 SIMT kernel keeps K / V of a head in shared memory
            o = torch.empty(B * T, H * 64, dtype=torch.float32, device=DEV)
            call("rtdf_attention", P(dev(qkv)), P(o), B, T, H, 0, 1, stream())
            d = float((o.cpu() - ref32).abs().max())
            out[f"B{B}_T{T}_H{H}_f32"] = d
            assert d <= 2e-5, out
        qb = qkv.to(torch.bfloat16)
        ref16 = _attn_ref(qb, B, T, H)
        for impl in impls:
            go = Guarded((B * T, H * 64), torch.bfloat16)
            o16 = go.t
            call("rtdf_attention", P(dev(qb)), P(o16), B, T, H, 1, impl, stream())
            go.check(f"rtdf_attention impl {impl} B{B} T{T} H{H}")
            d = float((o16.float().cpu() - ref16).abs().max())
            out[f"B{B}_T{T}_H{H}_bf16_impl{impl}"] = d
            assert d <= 0.03, out     # bf16 P and bf16 output rounding, |out| <~ 3
    return out


def check_graph_pool():
    from oracle.aasist_ref import GraphPool
    out = {}
    torch.manual_seed(7)
    for B, n, D in ((4, 42, 64), (3, 67, 64), (2, 33, 32), (5, 21, 32), (2, 1, 32), (2, 16, 32), (2, 123, 64), (2, 200, 32)):
        gp = GraphPool(0.5, D).eval()
        h = torch.randn(B, n, D)
        with torch.no_grad():
            ref, idx = gp(h, return_idx=True)
        k = ref.shape[1]
        o = torch.empty(B, k, D, dtype=torch.float32, device=DEV)
        io = torch.empty(B, k, dtype=torch.int32, device=DEV)
        call("rtdf_graph_pool", P(dev(h)), B, n, D, P(dev(gp.proj.weight.detach())), P(dev(gp.proj.bias.detach())),
             k, P(o), P(io), stream())
        same = bool((io.cpu().long() == idx).all())
        d = float((o.cpu() - ref).abs().max())
        out[f"B{B}_n{n}_D{D}"] = (same, d)
        assert same, out            # bit-exact node selection and order (distinct scores)
        assert d <= 1e-6, out
    # ties: equal scores keep the lower index first (documented tie rule)
    h = torch.ones(1, 8, 32)
    o = torch.empty(1, 4, 32, device=DEV)
    io = torch.empty(1, 4, dtype=torch.int32, device=DEV)
    w = torch.ones(32)
    call("rtdf_graph_pool", P(dev(h)), 1, 8, 32, P(dev(w)), P(dev(torch.zeros(1))), 4, P(o), P(io), stream())
    assert io.cpu().tolist() == [[0, 1, 2, 3]], io
    return out


def _gat_weights(nat, keep, att, att_b, a11, a22, a12, with_l, without_l, bn, temp):
    """ctypes rtdf_gat_weights from oracle sub-modules; `keep` holds the device tensors alive."""
    def d(t):
        t = dev(t.detach().float())
        keep.append(t)
        return t.data_ptr()
    w = nat.GatWeights()
    w.att_w, w.att_b, w.a11 = d(att), d(att_b), d(a11.reshape(-1))
    w.a22 = d(a22.reshape(-1)) if a22 is not None else None
    w.a12 = d(a12.reshape(-1)) if a12 is not None else None
    w.with_t, w.with_b = d(with_l.weight.t().contiguous()), d(with_l.bias)
    w.without_t, w.without_b = d(without_l.weight.t().contiguous()), d(without_l.bias)
    if bn is not None:
        s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        w.bn_s, w.bn_t = d(s), d(bn.bias - bn.running_mean * s)
    w.inv_temp = 1.0 / temp
    return w


def check_gat_rows(impl=0):
    """GraphAttentionLayer and HtrgGraphAttentionLayer rows (after the type projections) against the oracle."""
    import ctypes
    from oracle.aasist_ref import GraphAttentionLayer, HtrgGraphAttentionLayer, perturb_norm_stats
    from tests.util import native
    nat = native()
    out = {}
    torch.manual_seed(3)
    tol = 2e-5 if impl == 1 else 2e-4
    for B, n, D, DO, temp in ((3, 42, 64, 64, 2.0), (2, 66, 64, 64, 2.0), (2, 16, 64, 64, 2.0), (2, 1, 64, 64, 2.0),
                              (2, 17, 64, 64, 0.5), (2, 123, 64, 64, 2.0), (1, 200, 64, 64, 2.0)):
        layer = GraphAttentionLayer(D, DO, temperature=temp).eval()
        perturb_norm_stats(layer, seed=5)
        x = torch.randn(B, n, D)
        with torch.no_grad():
            ref = layer(x)
        keep = []
        w = _gat_weights(nat, keep, layer.att_proj.weight, layer.att_proj.bias, layer.att_weight, None, None,
                         layer.proj_with_att, layer.proj_without_att, layer.bn, temp)
        o = torch.empty(B, n, DO, device=DEV)
        call("rtdf_gat_rows", D, DO, P(dev(x)), B, n, n, ctypes.byref(w), P(o), None, None, None, impl, stream())
        d = float((o.cpu() - ref).abs().max())
        out[f"gat_B{B}_n{n}"] = d
        assert d <= tol, out
    for B, n1, n2, D, DO, temp in ((3, 33, 21, 64, 32, 100.0), (2, 16, 10, 32, 32, 100.0), (2, 5, 3, 64, 32, 1.0), (2, 61, 21, 64, 32, 100.0),
                                   (2, 1, 1, 32, 32, 100.0)):
        layer = HtrgGraphAttentionLayer(D, DO, temperature=temp).eval()
        perturb_norm_stats(layer, seed=6)
        x1, x2, master = torch.randn(B, n1, D), torch.randn(B, n2, D), torch.randn(B, 1, D)
        with torch.no_grad():
            r1, r2, rm = layer(x1, x2, master=master)
            xcat = torch.cat([layer.proj_type1(x1), layer.proj_type2(x2)], dim=1)   # type_proj_kernel's output
        keep = []
        w = _gat_weights(nat, keep, layer.att_proj.weight, layer.att_proj.bias, layer.att_weight11, layer.att_weight22,
                         layer.att_weight12, layer.proj_with_att, layer.proj_without_att, layer.bn, temp)
        wm = _gat_weights(nat, keep, layer.att_projM.weight, layer.att_projM.bias, layer.att_weightM, None, None,
                          layer.proj_with_attM, layer.proj_without_attM, None, temp)
        n = n1 + n2
        o = torch.empty(B, n, DO, device=DEV)
        mo = torch.empty(B, DO, device=DEV)
        call("rtdf_gat_rows", D, DO, P(dev(xcat)), B, n, n1, ctypes.byref(w), P(o), P(dev(master.reshape(B, D))),
             ctypes.byref(wm), P(mo), impl, stream())
        d = max(float((o.cpu() - torch.cat([r1, r2], 1)).abs().max()), float((mo.cpu() - rm.reshape(B, DO)).abs().max()))
        out[f"hsgal_B{B}_n{n1}+{n2}_D{D}"] = d
        assert d <= tol, out
    return out


def check_conformer_attention(is_bf16=1, impl=0, shapes=((2, 50, 4, 36), (2, 200, 4, 36), (1, 202, 4, 36), (2, 17, 2, 36))):
    """Shaw relative-position MHSA of lucidrains' ConformerBlock (instantiated at reference
    models/conformer_baseline.py:16-18; oracle/conformer_block_ref.py Attention) from a packed [q|k|v] matrix."""
    from oracle.conformer_block_ref import Attention
    g = torch.Generator().manual_seed(21)
    out = {}
    for B, n, heads, dh in shapes:
        E = heads * dh
        att = Attention(E, heads=heads, dim_head=dh).eval()
        with torch.no_grad():
            att.rel_pos_emb.weight.copy_(torch.randn(1025, dh, generator=g) * 0.5)
        qkv = torch.randn(B * n, 3 * E, generator=g)
        if is_bf16:
            qkv = qkv.to(torch.bfloat16).float()
        q, k, v = (qkv[:, i * E:(i + 1) * E].reshape(B, n, heads, dh).transpose(1, 2) for i in range(3))
        dots = torch.einsum("bhid,bhjd->bhij", q, k) * att.scale
        seq = torch.arange(n)
        dist = (seq[:, None] - seq[None, :]).clamp(-512, 512) + 512
        rel = att.rel_pos_emb.weight.detach()[dist]
        dots = dots + torch.einsum("bhnd,nrd->bhnr", q, rel) * att.scale
        ref = torch.einsum("bhij,bhjd->bhid", dots.softmax(-1), v).transpose(1, 2).reshape(B * n, E)
        dt = torch.bfloat16 if is_bf16 else torch.float32
        o = torch.zeros(B * n, E, dtype=dt, device=DEV)
        call("rtdf_conformer_attention", P(dev(qkv.to(dt))), P(dev(att.rel_pos_emb.weight.detach())), P(o), B, n, heads, dh,
             is_bf16, impl, stream())
        d = float((o.float().cpu() - ref).abs().max())
        out[f"B{B}_n{n}_h{heads}_bf16{is_bf16}_impl{impl}"] = d
        assert d <= (0.03 if is_bf16 else 2e-5), out     # bf16: rel-pos table / P / output rounded to bf16, |out| <~ 3
    return out


def check_conformer_glu_dwconv(is_bf16=1, shapes=((2, 50, 288, 31), (2, 200, 288, 31), (1, 313, 288, 31), (3, 7, 288, 31),
                                                   (2, 64, 64, 15))):
    """GLU -> depth-wise conv ("same" padding) -> BatchNorm1d (eval) -> Swish of the Conformer convolution module
    (oracle/conformer_block_ref.py ConformerConvModule.net[3:7])."""
    from oracle.conformer_block_ref import ConformerConvModule
    g = torch.Generator().manual_seed(22)
    out = {}
    for B, n, inner, k in shapes:
        mod = ConformerConvModule(inner // 2, expansion_factor=2, kernel_size=k).eval()
        glu, dw, bn, swish = mod.net[3], mod.net[4], mod.net[5], mod.net[6]
        with torch.no_grad():
            bn.running_mean.copy_(torch.randn(inner, generator=g) * 0.1)
            bn.running_var.copy_(torch.rand(inner, generator=g) + 0.5)
            bn.weight.copy_(1 + 0.1 * torch.randn(inner, generator=g))
            bn.bias.copy_(0.1 * torch.randn(inner, generator=g))
        x = torch.randn(B * n, 2 * inner, generator=g)
        if is_bf16:
            x = x.to(torch.bfloat16).float()
        with torch.no_grad():
            ref = swish(bn(dw(glu(x.reshape(B, n, 2 * inner).transpose(1, 2))))).transpose(1, 2).reshape(B * n, inner)
        s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        t = bn.bias - bn.running_mean * s
        dt = torch.bfloat16 if is_bf16 else torch.float32
        o = torch.zeros(B * n, inner, dtype=dt, device=DEV)
        call("rtdf_conformer_glu_dwconv", P(dev(x.to(dt))), P(o), B, n, inner, k, P(dev(dw.conv.weight.detach().reshape(inner, k))),
             P(dev(dw.conv.bias.detach())), P(dev(s.detach())), P(dev(t.detach())), is_bf16, stream())
        d = float((o.float().cpu() - ref).abs().max())
        out[f"B{B}_n{n}_c{inner}_k{k}_bf16{is_bf16}"] = d
        assert d <= (0.03 if is_bf16 else 2e-5), out
    return out
