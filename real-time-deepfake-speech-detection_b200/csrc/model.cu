// Context lifecycle, weight packing and the scoring forward (waveform -> XLS-R -> back-end -> logits).
#include "model.cuh"

#include <stdlib.h>

#include <iterator>
#include <set>

#include "attention.cuh"
#include "conformer.cuh"
#include "conv_tc.cuh"
#include "frontend.cuh"
#include "gemm_simt.cuh"

using namespace rtdf;

namespace rtdf {

static const char* kSsl = "ssl_model.model.";

// ---------------------------------------------------------------------------------------------
// allocation / lookup helpers
// ---------------------------------------------------------------------------------------------
template <typename T>
static int dalloc(rtdf_ctx* c, long long n, T** out) {
  void* p = nullptr;
  RTDF_CHECK_CUDA(cudaMalloc(&p, (size_t)n * sizeof(T)));
  c->owned.push_back(p);
  *out = static_cast<T*>(p);
  return RTDF_OK;
}

static int get_raw(rtdf_ctx* c, const std::string& key, const Raw** out, long long expect_numel = -1) {
  auto it = c->raw.find(key);
  if (it == c->raw.end()) {
    set_error("rtdf_finalize: missing state-dict key '%s'", key.c_str());
    return RTDF_ERR_STATE;
  }
  if (expect_numel >= 0 && it->second.numel != expect_numel) {
    set_error("rtdf_finalize: key '%s' has %lld elements, expected %lld", key.c_str(), it->second.numel, expect_numel);
    return RTDF_ERR_INVALID;
  }
  *out = &it->second;
  return RTDF_OK;
}

static int get_ptr(rtdf_ctx* c, const std::string& key, const float** out, long long expect_numel) {
  const Raw* r;
  RTDF_TRY(get_raw(c, key, &r, expect_numel));
  *out = r->p;
  return RTDF_OK;
}

// bf16 mode keeps one packed bf16 copy of every GEMM-sized weight; its fp32 source (the state-dict copy made by
// rtdf_load_weight, or an fp32 staging buffer of the packing) is released at the end of rtdf_finalize.
constexpr long long kDropMinNumel = 32768;
static void drop_after_finalize(rtdf_ctx* c, const void* p) {
  if (c->d.precision == RTDF_PREC_BF16 && p) c->scratch.push_back(const_cast<void*>(p));
}

static int to_bf16(rtdf_ctx* c, const float* src, long long n, const bf16** out) {
  bf16* p;
  RTDF_TRY(dalloc(c, n, &p));
  RTDF_TRY(cast_f32_to_bf16(0, src, p, n));
  *out = p;
  return RTDF_OK;
}

static int make_lin(rtdf_ctx* c, const std::string& prefix, int n, int k, bool bias, Lin* out) {
  out->n = n;
  out->k = k;
  RTDF_TRY(get_ptr(c, prefix + ".weight", &out->w, (long long)n * k));
  if (bias) RTDF_TRY(get_ptr(c, prefix + ".bias", &out->b, n));
  if (c->d.precision == RTDF_PREC_BF16) {
    RTDF_TRY(to_bf16(c, out->w, (long long)n * k, &out->wb));
    if ((long long)n * k >= kDropMinNumel) {   // small matrices (fc5, out_layer) are also read as fp32 by row kernels
      drop_after_finalize(c, out->w);
      out->w = nullptr;
    }
  }
  return RTDF_OK;
}

static int make_ln(rtdf_ctx* c, const std::string& prefix, int n, Norm* out) {
  RTDF_TRY(get_ptr(c, prefix + ".weight", &out->g, n));
  RTDF_TRY(get_ptr(c, prefix + ".bias", &out->b, n));
  return RTDF_OK;
}

static int make_bn(rtdf_ctx* c, const std::string& prefix, int n, Norm* out) {
  const float *w, *b, *m, *v;
  RTDF_TRY(get_ptr(c, prefix + ".weight", &w, n));
  RTDF_TRY(get_ptr(c, prefix + ".bias", &b, n));
  RTDF_TRY(get_ptr(c, prefix + ".running_mean", &m, n));
  RTDF_TRY(get_ptr(c, prefix + ".running_var", &v, n));
  float *s, *t;
  RTDF_TRY(dalloc(c, n, &s));
  RTDF_TRY(dalloc(c, n, &t));
  RTDF_TRY(bn_fold(0, w, b, m, v, 1e-5f, s, t, n));
  out->g = s;
  out->b = t;
  return RTDF_OK;
}

static int make_transposed(rtdf_ctx* c, const std::string& key, int r, int cols, const float** out) {
  const float* src;
  RTDF_TRY(get_ptr(c, key, &src, (long long)r * cols));
  float* dst;
  RTDF_TRY(dalloc(c, (long long)r * cols, &dst));
  RTDF_TRY(transpose_f32(0, src, dst, r, cols));
  *out = dst;
  return RTDF_OK;
}

static int scalar_bn(rtdf_ctx* c, const std::string& prefix, float* s, float* t) {
  Norm n;
  RTDF_TRY(make_bn(c, prefix, 1, &n));
  RTDF_CHECK_CUDA(cudaDeviceSynchronize());
  RTDF_CHECK_CUDA(cudaMemcpy(s, n.g, sizeof(float), cudaMemcpyDeviceToHost));
  RTDF_CHECK_CUDA(cudaMemcpy(t, n.b, sizeof(float), cudaMemcpyDeviceToHost));
  return RTDF_OK;
}

// ---------------------------------------------------------------------------------------------
// packing
// ---------------------------------------------------------------------------------------------
static int pack_xlsr(rtdf_ctx* c) {
  const std::string P = kSsl;
  const bool bf = c->d.precision == RTDF_PREC_BF16;
  static const int ks[7] = {10, 3, 3, 3, 3, 2, 2};
  static const int ss[7] = {5, 2, 2, 2, 2, 2, 2};
  // fairseq extractor_mode: "layer_norm" (XLS-R: conv bias + per-frame LayerNorm(512) after every conv, keys
  // conv_layers.{i}.2.1.*) or "default" (wav2vec2-base style: GroupNorm(512, 512) over time after conv-0 only, keys
  // conv_layers.0.2.*, usually no conv bias) -- recognised from the state-dict keys.
  const std::string fe0 = P + "feature_extractor.conv_layers.0";
  c->fe_group_norm = c->raw.count(fe0 + ".2.1.weight") == 0 && c->raw.count(fe0 + ".2.weight") != 0;
  for (int i = 0; i < 7; ++i) {
    FeConv& f = c->fe[i];
    f.k = ks[i];
    f.stride = ss[i];
    const std::string cp = P + "feature_extractor.conv_layers." + std::to_string(i);
    if (c->fe_group_norm) {
      if (i == 0) RTDF_TRY(make_ln(c, cp + ".2", 512, &f.ln));      // GroupNorm affine (per channel)
    } else {
      if (c->raw.count(cp + ".2.1.weight") == 0) {
        set_error("rtdf_finalize: neither '%s.2.1.weight' (extractor_mode=layer_norm) nor '%s.2.weight' (group norm) found",
                  cp.c_str(), fe0.c_str());
        return RTDF_ERR_STATE;
      }
      RTDF_TRY(make_ln(c, cp + ".2.1", 512, &f.ln));
    }
    if (c->raw.count(cp + ".0.bias")) RTDF_TRY(get_ptr(c, cp + ".0.bias", &f.lin.b, 512));   // conv_bias=False: no key
    else if (!c->fe_group_norm) RTDF_TRY(get_ptr(c, cp + ".0.bias", &f.lin.b, 512));          // reports the missing key
    const int ci = i == 0 ? 1 : 512;
    const float* w;
    RTDF_TRY(get_ptr(c, cp + ".0.weight", &w, 512LL * ci * f.k));
    float* wp;
    RTDF_TRY(dalloc(c, 512LL * ci * f.k, &wp));
    if (i == 0) {
      RTDF_TRY(transpose_f32(0, w, wp, 512, 10));  // [512][10] -> [10][512]
      if (bf && !c->fe_group_norm) {               // hi/lo-split K = 32 operand of the tensor-core conv-0
        bf16* w0;
        RTDF_TRY(dalloc(c, 512LL * 32, &w0));
        RTDF_TRY(conv0_tc_pack_weight(0, w, w0));
        c->conv0_tc_w = w0;
      }
    } else {
      RTDF_TRY(permute_conv_weight(0, w, wp, 512, 512, f.k));  // [co][ci][k] -> [co][k][ci]
    }
    f.lin.w = wp;
    f.lin.n = 512;
    f.lin.k = ci * f.k;
    if (bf && i > 0) {
      RTDF_TRY(to_bf16(c, wp, 512LL * ci * f.k, &f.lin.wb));
      drop_after_finalize(c, wp);
      drop_after_finalize(c, w);
      f.lin.w = nullptr;
    }
  }
  RTDF_TRY(make_ln(c, P + "layer_norm", 512, &c->fp_ln));
  RTDF_TRY(make_lin(c, P + "post_extract_proj", 1024, 512, true, &c->proj));
  {  // positional conv: weight-norm fold + [co][k][ci] layout
    const float *v, *g;
    RTDF_TRY(get_ptr(c, P + "encoder.pos_conv.0.weight_v", &v, 1024LL * 64 * 128));
    RTDF_TRY(get_ptr(c, P + "encoder.pos_conv.0.weight_g", &g, 128));
    RTDF_TRY(get_ptr(c, P + "encoder.pos_conv.0.bias", &c->pos.b, 1024));
    float* wp;
    RTDF_TRY(dalloc(c, 1024LL * 8192, &wp));
    RTDF_TRY(posconv_fold_weight(0, v, g, wp, 1024, 64, 128));
    c->pos.w = wp;
    c->pos.n = 1024;
    c->pos.k = 8192;
    if (bf) {
      RTDF_TRY(to_bf16(c, wp, 1024LL * 8192, &c->pos.wb));
      drop_after_finalize(c, wp);
      drop_after_finalize(c, v);
      c->pos.w = nullptr;
    }
  }
  c->layers.resize(c->d.n_layers);
  for (int l = 0; l < c->d.n_layers; ++l) {
    XlsrLayer& L = c->layers[l];
    const std::string lp = P + "encoder.layers." + std::to_string(l);
    // fused QKV with the 1/sqrt(64) query scale folded in (exact: power of two)
    float *w, *b;
    RTDF_TRY(dalloc(c, 3072LL * 1024, &w));
    RTDF_TRY(dalloc(c, 3072, &b));
    const char* names[3] = {"q_proj", "k_proj", "v_proj"};
    for (int j = 0; j < 3; ++j) {
      const float *wj, *bj;
      RTDF_TRY(get_ptr(c, lp + ".self_attn." + names[j] + ".weight", &wj, 1024LL * 1024));
      RTDF_TRY(get_ptr(c, lp + ".self_attn." + names[j] + ".bias", &bj, 1024));
      RTDF_CHECK_CUDA(cudaMemcpy(w + j * 1024LL * 1024, wj, 1024LL * 1024 * 4, cudaMemcpyDeviceToDevice));
      RTDF_CHECK_CUDA(cudaMemcpy(b + j * 1024, bj, 1024 * 4, cudaMemcpyDeviceToDevice));
      drop_after_finalize(c, wj);
    }
    RTDF_TRY(scale_rows_f32(0, w, 1024, 1024, 0.125f));
    RTDF_TRY(scale_rows_f32(0, b, 1, 1024, 0.125f));
    L.qkv.w = w;
    L.qkv.b = b;
    L.qkv.n = 3072;
    L.qkv.k = 1024;
    RTDF_TRY(make_ln(c, lp + ".self_attn_layer_norm", 1024, &L.ln1));
    RTDF_TRY(make_ln(c, lp + ".final_layer_norm", 1024, &L.ln2));
    RTDF_TRY(make_lin(c, lp + ".self_attn.out_proj", 1024, 1024, true, &L.out));
    RTDF_TRY(make_lin(c, lp + ".fc2", 1024, 4096, true, &L.fc2));
    if (bf && !c->ln_fold) {
      RTDF_TRY(to_bf16(c, w, 3072LL * 1024, &L.qkv.wb));
      drop_after_finalize(c, w);
      L.qkv.w = nullptr;
      RTDF_TRY(make_lin(c, lp + ".fc1", 4096, 1024, true, &L.fc1));
    } else if (bf) {
      // RTDF_LN_FOLD=1: the LayerNorm in front of the QKV / fc1 projections is folded into their weights (see run_frontend)
      auto fold = [&](const float* wf, const float* bias, const Norm& ln, int n, Lin* lin, const float** cv, const float** dv) -> int {
        bf16* wb;
        float *cc, *dd;
        RTDF_TRY(dalloc(c, (long long)n * 1024, &wb));
        RTDF_TRY(dalloc(c, n, &cc));
        RTDF_TRY(dalloc(c, n, &dd));
        RTDF_TRY(fold_ln_weight(0, wf, ln.g, ln.b, bias, n, 1024, wb, cc, dd));
        lin->wb = wb;
        lin->w = nullptr;
        *cv = cc;
        *dv = dd;
        return RTDF_OK;
      };
      RTDF_TRY(fold(w, b, L.ln1, 3072, &L.qkv, &L.qkv_c, &L.qkv_d));
      drop_after_finalize(c, w);
      const float *w1, *b1;
      RTDF_TRY(get_ptr(c, lp + ".fc1.weight", &w1, 4096LL * 1024));
      RTDF_TRY(get_ptr(c, lp + ".fc1.bias", &b1, 4096));
      L.fc1.n = 4096;
      L.fc1.k = 1024;
      L.fc1.b = b1;
      RTDF_TRY(fold(w1, b1, L.ln2, 4096, &L.fc1, &L.fc1_c, &L.fc1_d));
      drop_after_finalize(c, w1);
    } else {
      RTDF_TRY(make_lin(c, lp + ".fc1", 4096, 1024, true, &L.fc1));
    }
  }
  RTDF_TRY(make_ln(c, P + "encoder.layer_norm", 1024, &c->enc_ln));
  return RTDF_OK;
}

static int pack_gat(rtdf_ctx* c, const std::string& p, int d, int dout, float temp, GatRowWeights* w) {
  RTDF_TRY(get_ptr(c, p + ".att_proj.weight", &w->att_w, (long long)dout * d));
  RTDF_TRY(get_ptr(c, p + ".att_proj.bias", &w->att_b, dout));
  RTDF_TRY(get_ptr(c, p + ".att_weight", &w->a11, dout));
  RTDF_TRY(make_transposed(c, p + ".proj_with_att.weight", dout, d, &w->with_t));
  RTDF_TRY(get_ptr(c, p + ".proj_with_att.bias", &w->with_b, dout));
  RTDF_TRY(make_transposed(c, p + ".proj_without_att.weight", dout, d, &w->without_t));
  RTDF_TRY(get_ptr(c, p + ".proj_without_att.bias", &w->without_b, dout));
  Norm bn;
  RTDF_TRY(make_bn(c, p + ".bn", dout, &bn));
  w->bn_s = bn.g;
  w->bn_t = bn.b;
  w->inv_temp = 1.0f / temp;
  return RTDF_OK;
}

static int pack_hsgal(rtdf_ctx* c, const std::string& p, int d, int dout, float temp, HsGalW* h) {
  h->d = d;
  h->dout = dout;
  RTDF_TRY(make_transposed(c, p + ".proj_type1.weight", d, d, &h->t1_wt));
  RTDF_TRY(get_ptr(c, p + ".proj_type1.bias", &h->t1_b, d));
  RTDF_TRY(make_transposed(c, p + ".proj_type2.weight", d, d, &h->t2_wt));
  RTDF_TRY(get_ptr(c, p + ".proj_type2.bias", &h->t2_b, d));
  GatRowWeights& r = h->rows;
  RTDF_TRY(get_ptr(c, p + ".att_proj.weight", &r.att_w, (long long)dout * d));
  RTDF_TRY(get_ptr(c, p + ".att_proj.bias", &r.att_b, dout));
  RTDF_TRY(get_ptr(c, p + ".att_weight11", &r.a11, dout));
  RTDF_TRY(get_ptr(c, p + ".att_weight22", &r.a22, dout));
  RTDF_TRY(get_ptr(c, p + ".att_weight12", &r.a12, dout));
  RTDF_TRY(make_transposed(c, p + ".proj_with_att.weight", dout, d, &r.with_t));
  RTDF_TRY(get_ptr(c, p + ".proj_with_att.bias", &r.with_b, dout));
  RTDF_TRY(make_transposed(c, p + ".proj_without_att.weight", dout, d, &r.without_t));
  RTDF_TRY(get_ptr(c, p + ".proj_without_att.bias", &r.without_b, dout));
  Norm bn;
  RTDF_TRY(make_bn(c, p + ".bn", dout, &bn));
  r.bn_s = bn.g;
  r.bn_t = bn.b;
  r.inv_temp = 1.0f / temp;
  GatRowWeights& m = h->master;
  RTDF_TRY(get_ptr(c, p + ".att_projM.weight", &m.att_w, (long long)dout * d));
  RTDF_TRY(get_ptr(c, p + ".att_projM.bias", &m.att_b, dout));
  RTDF_TRY(get_ptr(c, p + ".att_weightM", &m.a11, dout));
  RTDF_TRY(make_transposed(c, p + ".proj_with_attM.weight", dout, d, &m.with_t));
  RTDF_TRY(get_ptr(c, p + ".proj_with_attM.bias", &m.with_b, dout));
  RTDF_TRY(make_transposed(c, p + ".proj_without_attM.weight", dout, d, &m.without_t));
  RTDF_TRY(get_ptr(c, p + ".proj_without_attM.bias", &m.without_b, dout));
  m.inv_temp = 1.0f / temp;
  return RTDF_OK;
}

static int pack_pool(rtdf_ctx* c, const std::string& p, int d, PoolW* w) {
  RTDF_TRY(get_ptr(c, p + ".proj.weight", &w->w, d));
  RTDF_TRY(get_ptr(c, p + ".proj.bias", &w->b, 1));
  return RTDF_OK;
}

static bool aasist_tc(const rtdf_ctx* c) {
  return c->d.precision == RTDF_PREC_BF16 && c->d.aasist_conv_impl != 2;
}

// (hi, lo) bf16 chunks [n_chunks][co][kw] of a conv / linear weight for the shifted-row tcgen05 kernel
static int pack_tc_conv(rtdf_ctx* c, const float* src, int co, int ci, int taps /*KH*3, or 0 = linear*/, TcConvW* out) {
  const int kw = ci < 64 ? ci : 64;
  long long off[kConvTcMaxChunks];
  long long stride_o, stride_k;
  int n_chunks;
  if (taps > 0) {          // conv weight [co][ci][kh][3]: chunk = tap, k = input channel
    RTDF_REQUIRE(ci <= 64, "pack_tc_conv: conv with more than 64 input channels");
    n_chunks = taps;
    stride_o = (long long)ci * taps;
    stride_k = taps;
    for (int t = 0; t < taps; ++t) off[t] = t;
  } else {                 // linear weight [co][ci]: chunk = 64-channel block
    n_chunks = (ci + 63) / 64;
    stride_o = ci;
    stride_k = 1;
    for (int t = 0; t < n_chunks; ++t) off[t] = 64LL * t;
  }
  bf16 *hi, *lo;
  RTDF_TRY(dalloc(c, (long long)n_chunks * co * kw, &hi));
  RTDF_TRY(dalloc(c, (long long)n_chunks * co * kw, &lo));
  RTDF_TRY(conv_tc_pack_weight(0, src, co, kw, n_chunks, stride_o, stride_k, off, hi, lo));
  out->hi = hi;
  out->lo = lo;
  out->ci = ci;
  out->co = co;
  out->n_chunks = n_chunks;
  return RTDF_OK;
}

static int pack_aasist(rtdf_ctx* c) {
  AasistW& a = c->aasist;
  RTDF_TRY(make_lin(c, "LL", 128, 1024, true, &a.LL));
  RTDF_TRY(scalar_bn(c, "first_bn", &a.first_bn_s, &a.first_bn_t));
  static const int filt[6][2] = {{1, 32}, {32, 32}, {32, 64}, {64, 64}, {64, 64}, {64, 64}};
  for (int i = 0; i < 6; ++i) {
    ResBlockW& b = a.blocks[i];
    b.ci = filt[i][0];
    b.co = filt[i][1];
    const std::string p = "encoder." + std::to_string(i) + ".0";
    // conv weights [Co][Ci][KH][3] -> [Ci][KH][3][Co]
    RTDF_TRY(make_transposed(c, p + ".conv1.weight", b.co, b.ci * 6, &b.conv1_w));
    RTDF_TRY(get_ptr(c, p + ".conv1.bias", &b.conv1_b, b.co));
    RTDF_TRY(make_bn(c, p + ".bn2", b.co, &b.bn2));
    RTDF_TRY(make_transposed(c, p + ".conv2.weight", b.co, b.co * 6, &b.conv2_w));
    RTDF_TRY(get_ptr(c, p + ".conv2.bias", &b.conv2_b, b.co));
    if (b.ci != b.co) {
      RTDF_TRY(make_transposed(c, p + ".conv_downsample.weight", b.co, b.ci * 3, &b.ds_w));
      RTDF_TRY(get_ptr(c, p + ".conv_downsample.bias", &b.ds_b, b.co));
    }
    // bn1.* exists in the state dict but its output is discarded by the reference (aasist_modules.py:376-383)
    if (aasist_tc(c)) {
      const float* w;
      RTDF_TRY(get_ptr(c, p + ".conv1.weight", &w, (long long)b.co * b.ci * 6));
      if (i == 0) b.conv1_raw = w;
      else RTDF_TRY(pack_tc_conv(c, w, b.co, b.ci, 6, &b.tc1));
      RTDF_TRY(get_ptr(c, p + ".conv2.weight", &w, (long long)b.co * b.co * 6));
      RTDF_TRY(pack_tc_conv(c, w, b.co, b.co, 6, &b.tc2));
      if (b.ci != b.co) {
        RTDF_TRY(get_ptr(c, p + ".conv_downsample.weight", &w, (long long)b.co * b.ci * 3));
        if (i == 0) b.ds_raw = w;
        else RTDF_TRY(pack_tc_conv(c, w, b.co, b.ci, 3, &b.tcd));
      }
    }
  }
  RTDF_TRY(make_bn(c, "first_bn1", 64, &a.first_bn1));
  RTDF_TRY(make_transposed(c, "attention.0.weight", 128, 64, &a.att_w1t));
  RTDF_TRY(get_ptr(c, "attention.0.bias", &a.att_b1, 128));
  RTDF_TRY(make_bn(c, "attention.2", 128, &a.att_bn));
  RTDF_TRY(make_transposed(c, "attention.3.weight", 64, 128, &a.att_w2t));
  RTDF_TRY(get_ptr(c, "attention.3.bias", &a.att_b2, 64));
  if (aasist_tc(c)) {
    // attention: conv1x1(64->128) -> SELU -> BN -> conv1x1(128->64).  The eval BatchNorm (scale s, shift t) folds
    // exactly into the second conv: W2' = W2 diag(s), b2' = b2 + W2 t.
    const float *w1, *w2, *b2;
    RTDF_TRY(get_ptr(c, "attention.0.weight", &w1, 128 * 64));
    RTDF_TRY(get_ptr(c, "attention.3.weight", &w2, 64 * 128));
    RTDF_TRY(get_ptr(c, "attention.3.bias", &b2, 64));
    RTDF_TRY(pack_tc_conv(c, w1, 128, 64, 0, &a.att1));
    RTDF_CHECK_CUDA(cudaDeviceSynchronize());
    std::vector<float> hw2(64 * 128), hb2(64), hs(128), ht(128);
    RTDF_CHECK_CUDA(cudaMemcpy(hw2.data(), w2, hw2.size() * 4, cudaMemcpyDeviceToHost));
    RTDF_CHECK_CUDA(cudaMemcpy(hb2.data(), b2, hb2.size() * 4, cudaMemcpyDeviceToHost));
    RTDF_CHECK_CUDA(cudaMemcpy(hs.data(), a.att_bn.g, 128 * 4, cudaMemcpyDeviceToHost));
    RTDF_CHECK_CUDA(cudaMemcpy(ht.data(), a.att_bn.b, 128 * 4, cudaMemcpyDeviceToHost));
    for (int o = 0; o < 64; ++o) {
      double acc = hb2[o];
      for (int j = 0; j < 128; ++j) {
        acc += (double)hw2[o * 128 + j] * ht[j];
        hw2[o * 128 + j] *= hs[j];
      }
      hb2[o] = (float)acc;
    }
    float *dw2, *db2;
    RTDF_TRY(dalloc(c, 64 * 128, &dw2));
    RTDF_TRY(dalloc(c, 64, &db2));
    RTDF_CHECK_CUDA(cudaMemcpy(dw2, hw2.data(), hw2.size() * 4, cudaMemcpyHostToDevice));
    RTDF_CHECK_CUDA(cudaMemcpy(db2, hb2.data(), hb2.size() * 4, cudaMemcpyHostToDevice));
    RTDF_TRY(pack_tc_conv(c, dw2, 64, 128, 0, &a.att2));
    a.att_b2_folded = db2;
  }
  RTDF_TRY(get_ptr(c, "pos_S", &a.pos_S, 42 * 64));
  RTDF_TRY(get_ptr(c, "master1", &a.master1, 64));
  RTDF_TRY(get_ptr(c, "master2", &a.master2, 64));
  RTDF_TRY(pack_gat(c, "GAT_layer_S", 64, 64, 2.0f, &a.gat_S));
  RTDF_TRY(pack_gat(c, "GAT_layer_T", 64, 64, 2.0f, &a.gat_T));
  RTDF_TRY(pack_hsgal(c, "HtrgGAT_layer_ST11", 64, 32, 100.0f, &a.st11));
  RTDF_TRY(pack_hsgal(c, "HtrgGAT_layer_ST12", 32, 32, 100.0f, &a.st12));
  RTDF_TRY(pack_hsgal(c, "HtrgGAT_layer_ST21", 64, 32, 100.0f, &a.st21));
  RTDF_TRY(pack_hsgal(c, "HtrgGAT_layer_ST22", 32, 32, 100.0f, &a.st22));
  RTDF_TRY(pack_pool(c, "pool_S", 64, &a.pool_S));
  RTDF_TRY(pack_pool(c, "pool_T", 64, &a.pool_T));
  RTDF_TRY(pack_pool(c, "pool_hS1", 32, &a.pool_hS1));
  RTDF_TRY(pack_pool(c, "pool_hT1", 32, &a.pool_hT1));
  RTDF_TRY(pack_pool(c, "pool_hS2", 32, &a.pool_hS2));
  RTDF_TRY(pack_pool(c, "pool_hT2", 32, &a.pool_hT2));
  RTDF_TRY(get_ptr(c, "out_layer.weight", &a.out_w, 2 * 160));
  RTDF_TRY(get_ptr(c, "out_layer.bias", &a.out_b, 2));
  return RTDF_OK;
}

static int pack_conformer(rtdf_ctx* c) {
  ConformerW& w = c->conf;
  const int E = c->d.conf_emb, Hh = c->d.conf_heads, K = c->d.conf_kernel, dh = E / Hh, inner = 2 * E;
  RTDF_REQUIRE(E % Hh == 0 && E % 8 == 0 && K % 2 == 1 && dh <= 64, "conformer: unsupported emb %d heads %d kernel %d", E, Hh, K);
  RTDF_TRY(make_lin(c, "LL", E, 1024, true, &w.LL));
  RTDF_TRY(scalar_bn(c, "first_bn", &w.first_bn_s, &w.first_bn_t));
  RTDF_TRY(get_ptr(c, "conformer.class_token", &w.class_token, E));
  RTDF_TRY(make_lin(c, "conformer.fc5", 2, E, true, &w.fc5));
  w.blocks.resize(c->d.conf_blocks);
  for (int i = 0; i < c->d.conf_blocks; ++i) {
    ConformerBlockW& b = w.blocks[i];
    const std::string p = "conformer.encoder_blocks." + std::to_string(i);
    RTDF_TRY(make_ln(c, p + ".ff1.fn.norm", E, &b.ff1_ln));
    RTDF_TRY(make_lin(c, p + ".ff1.fn.fn.net.0", 4 * E, E, true, &b.ff1_a));
    RTDF_TRY(make_lin(c, p + ".ff1.fn.fn.net.3", E, 4 * E, true, &b.ff1_b));
    RTDF_TRY(make_ln(c, p + ".ff2.fn.norm", E, &b.ff2_ln));
    RTDF_TRY(make_lin(c, p + ".ff2.fn.fn.net.0", 4 * E, E, true, &b.ff2_a));
    RTDF_TRY(make_lin(c, p + ".ff2.fn.fn.net.3", E, 4 * E, true, &b.ff2_b));
    RTDF_TRY(make_ln(c, p + ".attn.norm", E, &b.attn_ln));
    {  // fused [to_q ; to_kv]
      const float *wq, *wkv;
      RTDF_TRY(get_ptr(c, p + ".attn.fn.to_q.weight", &wq, (long long)E * E));
      RTDF_TRY(get_ptr(c, p + ".attn.fn.to_kv.weight", &wkv, 2LL * E * E));
      float* f;
      RTDF_TRY(dalloc(c, 3LL * E * E, &f));
      RTDF_CHECK_CUDA(cudaMemcpy(f, wq, (size_t)E * E * 4, cudaMemcpyDeviceToDevice));
      RTDF_CHECK_CUDA(cudaMemcpy(f + (long long)E * E, wkv, (size_t)2 * E * E * 4, cudaMemcpyDeviceToDevice));
      b.qkv.w = f;
      b.qkv.n = 3 * E;
      b.qkv.k = E;
      if (c->d.precision == RTDF_PREC_BF16) RTDF_TRY(to_bf16(c, f, 3LL * E * E, &b.qkv.wb));
    }
    RTDF_TRY(make_lin(c, p + ".attn.fn.to_out", E, E, true, &b.attn_out));
    RTDF_TRY(get_ptr(c, p + ".attn.fn.rel_pos_emb.weight", &b.rel_pos, 1025LL * dh));
    RTDF_TRY(make_ln(c, p + ".conv.net.0", E, &b.conv_ln));
    RTDF_TRY(make_lin(c, p + ".conv.net.2", 2 * inner, E, true, &b.pw1));       // Conv1d weight [576][144][1]
    RTDF_TRY(get_ptr(c, p + ".conv.net.4.conv.weight", &b.dw_w, (long long)inner * K));
    RTDF_TRY(get_ptr(c, p + ".conv.net.4.conv.bias", &b.dw_b, inner));
    RTDF_TRY(make_bn(c, p + ".conv.net.5", inner, &b.dw_bn));
    RTDF_TRY(make_lin(c, p + ".conv.net.7", E, inner, true, &b.pw2));
    RTDF_TRY(make_ln(c, p + ".post_norm", E, &b.post_ln));
  }
  return RTDF_OK;
}

// ---------------------------------------------------------------------------------------------
// workspace plan
// ---------------------------------------------------------------------------------------------
struct Bump {
  char* base = nullptr;
  size_t off = 0;
  template <typename T>
  T* take(long long n) {
    off = align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += (size_t)(n > 0 ? n : 1) * sizeof(T);
    return p;
  }
};

struct Dims {
  int B, N, L[7], T, M;
};

static int make_dims(int B, int N, Dims* d) {
  RTDF_REQUIRE(B >= 1 && B <= 4096, "batch %d out of range (1..4096)", B);
  d->B = B;
  d->N = N;
  static const int ks[7] = {10, 3, 3, 3, 3, 2, 2};
  static const int ss[7] = {5, 2, 2, 2, 2, 2, 2};
  int n = N;
  for (int i = 0; i < 7; ++i) {
    n = n >= ks[i] ? (n - ks[i]) / ss[i] + 1 : 0;
    d->L[i] = n;
  }
  d->T = d->L[6];
  RTDF_REQUIRE(d->T >= 1, "utterance of %d samples is too short for the 7-layer conv encoder (>= 400 needed)", N);
  d->M = B * d->T;
  return RTDF_OK;
}

constexpr long long kSkinnyRows = 512;   // B*T up to which GEMMs are treated as weight-streaming problems
constexpr long long kPdlRows = 1024;     // B*T up to which programmatic dependent launch is switched on
constexpr long long kSmallConvRows = 16384;   // conv layers with at most this many output rows run as 64-wide GEMM + LN

struct FrontWs {
  float* wav_pe;           // pre-emphasised copy
  void* actA;              // conv ping
  void* actB;              // conv pong
  void* featln;            // (M,512)
  float* x;                // (M,1024) fp32 residual stream
  void* xb;                // (M,1024) normalised activations (bf16 | fp32)
  void* qkv;               // (M,3072)
  void* attn;              // (M,1024)
  void* hbuf;              // (M,4096)
  float* feats;            // (M,1024) fp32
  float2* stats;           // (M, 8) per-row partial (sum, sum of squares) of the residual stream (folded LayerNorm, bf16 mode)
  float* partials;         // [<= 8][M][1024] K-split partial sums of out_proj / fc2 (streaming-chunk regime only)
  float* gn_ws;            // group-norm extractor mode: conv-0 partial statistics + per-utterance scale / shift
  float* conv_f32;         // (rows, 512) pre-LayerNorm conv output of the short conv layers (streaming-chunk regime only)
  long long conv_f32_rows; // its capacity in rows (0 = not planned)
};

static void plan_front(const rtdf_ctx* c, const Dims& d, bool need_pe, bool own_feats, Bump& b, FrontWs* w) {
  const size_t es = c->d.precision == RTDF_PREC_BF16 ? 2 : 4;
  w->wav_pe = need_pe ? b.take<float>((long long)d.B * d.N) : nullptr;
  // +1 row of slack: garbage rows of the last implicit-GEMM tile may touch one frame past the end
  w->actA = b.take<char>(((long long)d.B * d.L[0] + 1) * 512 * es);
  w->actB = b.take<char>(((long long)d.B * d.L[1] + 1) * 512 * es);
  w->featln = b.take<char>((long long)d.M * 512 * es);
  w->x = b.take<float>((long long)d.M * 1024);
  w->xb = b.take<char>((long long)d.M * 1024 * es);
  w->qkv = b.take<char>((long long)d.M * 3072 * es);
  w->attn = b.take<char>((long long)d.M * 1024 * es);
  w->hbuf = b.take<char>((long long)d.M * 4096 * es);
  w->feats = own_feats ? b.take<float>((long long)d.M * 1024) : nullptr;
  w->stats = b.take<float2>((long long)d.M * 8);
  w->gn_ws = c->fe_group_norm ? b.take<float>((long long)conv0_gn_workspace_floats(d.B, d.N)) : nullptr;
  w->partials = d.M <= kSkinnyRows ? b.take<float>(8LL * d.M * 1024) : nullptr;
  w->conv_f32_rows = d.M <= kSkinnyRows && (long long)d.B * d.L[1] <= kSmallConvRows ? (long long)d.B * d.L[1] : 0;
  w->conv_f32 = w->conv_f32_rows > 0 ? b.take<float>(w->conv_f32_rows * 512) : nullptr;
}

struct AasistWs {
  void* featsb;     // bf16 copy of feats (bf16 mode)
  float* z;         // (M,128)
  float* r[4];      // (B,64,43,Tp) rotating conv buffers
  float* wmap;
  // tcgen05 path: zero-padded channels-last planes [B*Hp*Wp][C]
  struct Plane { float* f; bf16* hi; bf16* lo; } pl[2];   // block input / output (64 ch)
  bf16 *y_hi, *y_lo;                                       // conv1 output (64 ch)
  float* idt;                                              // conv_downsample output (64 ch)
  bf16 *hid_hi, *hid_lo;                                   // attention hidden (128 ch)
  float *eS, *eT, *gS, *gT, *oS, *oT;
  struct Br { float *hx, *hy, *ma, *pS, *pT, *hx2, *hy2, *mb; } br[2];
  int *idxS, *idxT;
};

constexpr int kAasistHp = 44;   // plane rows per utterance: hp = h + 1 for 42-row tensors, hp = h for conv1's 43 rows

static void plan_aasist(const rtdf_ctx* c, int B, int T, Bump& b, AasistWs* w) {
  const int Tp = T / 3, kT = Tp / 2 > 0 ? Tp / 2 : 1, kT2 = kT / 2 > 0 ? kT / 2 : 1;
  const long long M = (long long)B * T;
  w->featsb = c->d.precision == RTDF_PREC_BF16 ? (void*)b.take<bf16>(M * 1024) : nullptr;
  w->z = b.take<float>(M * 128);
  if (aasist_tc(c)) {
    const long long rows = (long long)B * kAasistHp * (Tp + 2);
    w->r[0] = b.take<float>((long long)B * 42 * Tp);      // stem output, one channel
    w->r[1] = w->r[2] = w->r[3] = nullptr;
    for (int i = 0; i < 2; ++i) {
      w->pl[i].f = b.take<float>(rows * 64);
      w->pl[i].hi = b.take<bf16>(rows * 64);
      w->pl[i].lo = b.take<bf16>(rows * 64);
    }
    w->y_hi = b.take<bf16>(rows * 64);
    w->y_lo = b.take<bf16>(rows * 64);
    w->idt = b.take<float>(rows * 64);
    w->hid_hi = b.take<bf16>(rows * 128);
    w->hid_lo = b.take<bf16>(rows * 128);
    w->wmap = b.take<float>(rows * 64);
  } else {
    for (int i = 0; i < 4; ++i) w->r[i] = b.take<float>((long long)B * 64 * 43 * Tp);
    w->wmap = b.take<float>((long long)B * 64 * 42 * Tp);
  }
  w->eS = b.take<float>((long long)B * 42 * 64);
  w->eT = b.take<float>((long long)B * Tp * 64);
  w->gS = b.take<float>((long long)B * 42 * 64);
  w->gT = b.take<float>((long long)B * Tp * 64);
  w->oS = b.take<float>((long long)B * 21 * 64);
  w->oT = b.take<float>((long long)B * kT * 64);
  for (int i = 0; i < 2; ++i) {
    w->br[i].hx = b.take<float>((long long)B * (kT + 21) * 64);
    w->br[i].hy = b.take<float>((long long)B * (kT + 21) * 32);
    w->br[i].ma = b.take<float>((long long)B * 32);
    w->br[i].pS = b.take<float>((long long)B * 10 * 32);
    w->br[i].pT = b.take<float>((long long)B * kT2 * 32);
    w->br[i].hx2 = b.take<float>((long long)B * (kT2 + 10) * 32);
    w->br[i].hy2 = b.take<float>((long long)B * (kT2 + 10) * 32);
    w->br[i].mb = b.take<float>((long long)B * 32);
  }
  w->idxS = b.take<int>((long long)B * 21);
  w->idxT = b.take<int>((long long)B * kT);
}

// ---------------------------------------------------------------------------------------------
// XLS-R front-end
// ---------------------------------------------------------------------------------------------
static TcOperandA plainA(const void* p, long long rows, long long k) {
  TcOperandA a;
  a.ptr = static_cast<const bf16*>(p);
  a.k_extent = k;
  a.rows_per_batch = rows;
  a.batches = 1;
  a.row_stride = k;
  return a;
}
static SimtOperandA plainAf(const void* p, long long rows, long long k) {
  SimtOperandA a;
  a.ptr = static_cast<const float*>(p);
  a.k_extent = k;
  a.rows_per_batch = rows;
  a.batches = 1;
  a.row_stride = k;
  return a;
}

static bool gemm_2sm_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RTDF_GEMM_2SM");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

static bool posconv_slab_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RTDF_POSCONV_IMPL");
    v = (e && e[0] == '1') ? 0 : 1;   // 1 = the tap-shifted GEMM it replaced (A/B timing)
  }
  return v == 1;
}

static bool skinny_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RTDF_SKINNY_GEMM");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// Streaming-chunk regime (64-wide weight-streaming tiles, split-K, GEMM + LayerNorm-row convs): chosen from the number
// of rows in flight unless the caller pinned the throughput regime (rtdf_set_regime), in which case an utterance's
// scores do not depend on the batch it travels in (ragged last batches, multi-GPU shards).
static bool skinny_rows(const rtdf_ctx* c, long long rows) {
  return c->regime != RTDF_REGIME_THROUGHPUT && rows <= kSkinnyRows && skinny_enabled();
}

// conv-0: tcgen05 implicit GEMM (im2col + full-row LayerNorm tile) once there are enough frames to fill the machine
// twice over with 128-row tiles; streaming chunks keep the SIMT kernel (32-frame CTAs).  RTDF_CONV0_IMPL=1: SIMT always.
static bool conv0_tc_wanted(const rtdf_ctx* c, long long frames) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RTDF_CONV0_IMPL");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  // (the throughput regime never looks at the batch size: per-utterance results must not depend on it)
  return v == 0 && c->conv0_tc_w && (c->regime == RTDF_REGIME_THROUGHPUT || frames >= 2LL * kNumSMs * 128);
}

static bool tail_split_enabled() {
  static int v = -1;
  if (v < 0) {
    // opt-in (r02 experiment, measured SLOWER: 13.03 vs 12.0 ms per step, profiles/r02_tail_split_ab.txt -- the second
    // persistent kernel's ramp-up / drain costs more than the partial round it removes)
    const char* e = getenv("RTDF_TAIL_SPLIT");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// L2 policy for the A operand of the transformer projections (activations that are dead once the GEMM has read them)
static int a_hint_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RTDF_A_HINT");
    v = e ? atoi(e) : 0;
  }
  return v;
}

static bool zigzag_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RTDF_ZIGZAG");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

static bool conv_2sm_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RTDF_CONV_2SM");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}


// y = epilogue(A W^T): dispatch on the context precision
static int linear(const rtdf_ctx* c, cudaStream_t s, const void* A, long long rows, const Lin& L, const TcEpilogue& e,
                  bool wide_tiles = false) {
  if (c->d.precision == RTDF_PREC_BF16) {
    int variant = L.n >= 256 ? 256 : (L.n >= 128 ? 128 : 64);
    if (variant == 256 && L.n % 256 == 0 && rows >= 2048 && gemm_2sm_enabled()) variant = 2256;   // CTA-pair tiles
    if (skinny_rows(c, rows) && L.n % 64 == 0 && !wide_tiles) {
      // Streaming chunks (batch 1-8 x 49 frames, or one 4 s utterance): the GEMM is a weight-streaming problem, so
      // the tile count -- not the tile shape -- sets the time: 64-wide tiles give 4x the CTAs of the 256-wide ones
      // (the residual GEMMs of the transformer layers additionally split K, see run_frontend).
      // ... unless 64-wide tiles no longer fit one round of the persistent kernel (3 - 4 row tiles x a wide N): 128-wide
      // tiles then finish in one round instead of two.
      const long long row_tiles = (rows + 127) / 128;
      const int width = (row_tiles * (L.n / 64) > kNumSMs && L.n % 128 == 0) ? 128 : 64;
      return tc_gemm(s, plainA(A, rows, L.k), L.wb, L.n, L.k, TC_PLAIN, width, e);
    }
    if (variant == 2256 && !e.xb_out && !e.fold_stats && !e.rowln_counters && tail_split_enabled()) {
      // Wave quantisation: the persistent CTA-pair kernel walks ceil(tiles / 74) rounds of 256 x 256 tiles, and at the
      // timed batch the last round is 10 - 70 % empty (out_proj / fc2: 200 tiles = 2.7 rounds).  The row blocks that do
      // not fill a whole number of rounds go to a second launch with 128 x 128 single-CTA tiles (a quarter of the work
      // each): its partial last round costs a quarter of a pair-tile time.  Same arithmetic per output element
      // (K order unchanged), so the result is bit-identical to the one-launch version.
      const long long tiles_m = (rows + 255) / 256, tiles_n = L.n / 256, pairs = kNumSMs / 2;
      const long long total = tiles_m * tiles_n;
      const long long full_rounds = total / pairs;
      const long long m_split = full_rounds * pairs / tiles_n;            // row blocks of whole rounds
      if (full_rounds >= 1 && m_split >= 1 && m_split < tiles_m) {
        const long long tail_rows = rows - m_split * 256;
        const long long tail_tiles = ((tail_rows + 127) / 128) * ((L.n + 127) / 128);
        const double before = (double)((total + pairs - 1) / pairs);
        const double after = (double)((m_split * tiles_n + pairs - 1) / pairs) + 0.29 * (double)((tail_tiles + kNumSMs - 1) / kNumSMs);
        if (after < 0.97 * before) {
          const long long r0 = m_split * 256;
          TcEpilogue et = e;
          if (et.resid) et.resid += r0 * et.ldr;
          if (et.out_f32) et.out_f32 += r0 * et.ld_f32;
          if (et.out_bf16) et.out_bf16 += r0 * et.ld_bf16;
          et.profile_as_wide = true;
          const bf16* At = static_cast<const bf16*>(A) + r0 * L.k;
          if (e.reverse_tiles) RTDF_TRY(tc_gemm(s, plainA(At, tail_rows, L.k), L.wb, L.n, L.k, TC_PLAIN, 128, et));
          RTDF_TRY(tc_gemm(s, plainA(A, r0, L.k), L.wb, L.n, L.k, TC_PLAIN, 2256, e));
          if (!e.reverse_tiles) RTDF_TRY(tc_gemm(s, plainA(At, tail_rows, L.k), L.wb, L.n, L.k, TC_PLAIN, 128, et));
          return RTDF_OK;
        }
      }
    }
    return tc_gemm(s, plainA(A, rows, L.k), L.wb, L.n, L.k, TC_PLAIN, variant, e);
  }
  return simt_gemm_f32(s, plainAf(A, rows, L.k), L.w, L.n, L.k, e);
}

static int run_frontend(rtdf_ctx* c, cudaStream_t s, const float* wav, const Dims& d, int preemph_on, float coef,
                        const FrontWs& w, float* feats, float* layer_taps = nullptr) {
  const bool bf = c->d.precision == RTDF_PREC_BF16;
  const int B = d.B, T = d.T;
  const long long M = d.M;
  if (preemph_on) {
    RTDF_TRY(preemph(s, wav, w.wav_pe, B, d.N, coef));
    wav = w.wav_pe;
  }
  // conv-0 (+bias +LN +GELU | +GroupNorm over time +GELU) -> channels-last (B,L1,512)
  const bool gn = c->fe_group_norm;
  if (gn)
    RTDF_TRY(conv0_gn_gelu(s, wav, B, d.N, c->fe[0].lin.w, c->fe[0].lin.b, c->fe[0].ln.g, c->fe[0].ln.b, 1e-5f, w.gn_ws,
                           bf ? nullptr : static_cast<float*>(w.actA), bf ? static_cast<bf16*>(w.actA) : nullptr));
  else if (bf && conv0_tc_wanted(c, (long long)B * d.L[0]))   // actB (conv-1's output buffer) is free: im2col scratch
    RTDF_TRY(conv0_tc_ln_gelu(s, wav, B, d.N, c->conv0_tc_w, c->fe[0].lin.b, c->fe[0].ln.g, c->fe[0].ln.b, 1e-5f,
                              static_cast<bf16*>(w.actB), static_cast<bf16*>(w.actA)));
  else
    RTDF_TRY(conv0_ln_gelu(s, wav, B, d.N, c->fe[0].lin.w, c->fe[0].lin.b, c->fe[0].ln.g, c->fe[0].ln.b, 1e-5f,
                           bf ? nullptr : static_cast<float*>(w.actA), bf ? static_cast<bf16*>(w.actA) : nullptr));
  void* cur = w.actA;
  void* nxt = w.actB;
  for (int i = 1; i < 7; ++i) {
    const FeConv& f = c->fe[i];
    const int Lin_ = d.L[i - 1], Lout = d.L[i];
    if (gn) {
      // extractor_mode="default": conv (no LayerNorm) -> GELU, as an implicit GEMM with the activation in the epilogue
      TcEpilogue e;
      e.bias = f.lin.b;
      e.act = ACT_GELU;
      if (bf) {
        TcOperandA a;
        a.ptr = static_cast<const bf16*>(cur);
        a.k_extent = (long long)f.k * 512;
        a.rows_per_batch = Lout;
        a.batches = B;
        a.row_stride = (long long)f.stride * 512;
        a.batch_stride = (long long)Lin_ * 512;
        e.out_bf16 = static_cast<bf16*>(nxt);
        e.ld_bf16 = 512;
        RTDF_TRY(tc_gemm(s, a, f.lin.wb, 512, f.k * 512, TC_PLAIN, 256, e));
      } else {
        SimtOperandA a;
        a.ptr = static_cast<const float*>(cur);
        a.k_extent = (long long)f.k * 512;
        a.rows_per_batch = Lout;
        a.batches = B;
        a.row_stride = (long long)f.stride * 512;
        a.batch_stride = (long long)Lin_ * 512;
        e.out_f32 = static_cast<float*>(nxt);
        e.ld_f32 = 512;
        RTDF_TRY(simt_gemm_f32(s, a, f.lin.w, 512, f.k * 512, e));
      }
    } else if (bf && w.conv_f32 && (long long)B * Lout <= w.conv_f32_rows && skinny_rows(c, M)) {
      // Streaming chunks: a full-row (N = 512) tile leaves one CTA per 128 output rows, each streaming the whole weight
      // matrix.  64-wide tiles give 8x the CTAs; bias goes in the GEMM, LayerNorm + GELU in a row kernel (fp32 in between).
      TcOperandA a;
      a.ptr = static_cast<const bf16*>(cur);
      a.k_extent = (long long)f.k * 512;
      a.rows_per_batch = Lout;
      a.batches = B;
      a.row_stride = (long long)f.stride * 512;
      a.batch_stride = (long long)Lin_ * 512;
      TcEpilogue e;
      e.bias = f.lin.b;
      e.out_f32 = w.conv_f32;
      e.ld_f32 = 512;
      RTDF_TRY(tc_gemm(s, a, f.lin.wb, 512, f.k * 512, TC_PLAIN, 64, e));
      RTDF_TRY(layernorm_rows_f32(s, w.conv_f32, (long long)B * Lout, 512, f.ln.g, f.ln.b, 1e-5f, ACT_GELU, nullptr,
                                  static_cast<bf16*>(nxt)));
    } else if (bf) {
      TcOperandA a;
      a.ptr = static_cast<const bf16*>(cur);
      a.k_extent = (long long)f.k * 512;
      a.rows_per_batch = Lout;
      a.batches = B;
      a.row_stride = (long long)f.stride * 512;
      a.batch_stride = (long long)Lin_ * 512;
      TcEpilogue e;
      e.bias = f.lin.b;
      e.act = ACT_GELU;
      e.ln_gamma = f.ln.g;
      e.ln_beta = f.ln.b;
      e.out_bf16 = static_cast<bf16*>(nxt);
      e.ld_bf16 = 512;
      // pipelined two-pass LayerNorm tile.  The CTA pair halves the weight stream per output row: +5 % on conv-1
      // (561 vs 592 us at B = 64), a tie on conv-2 and a loss on the short layers, so it is used for long outputs only.
      const int variant = conv_2sm_enabled() && (long long)B * Lout >= 200000 ? 516 : 515;
      RTDF_TRY(tc_gemm(s, a, f.lin.wb, 512, f.k * 512, TC_PLAIN, variant, e));
    } else {
      SimtOperandA a;
      a.ptr = static_cast<const float*>(cur);
      a.k_extent = (long long)f.k * 512;
      a.rows_per_batch = Lout;
      a.batches = B;
      a.row_stride = (long long)f.stride * 512;
      a.batch_stride = (long long)Lin_ * 512;
      TcEpilogue e;
      e.bias = f.lin.b;
      e.out_f32 = static_cast<float*>(nxt);
      e.ld_f32 = 512;
      RTDF_TRY(simt_gemm_f32(s, a, f.lin.w, 512, f.k * 512, e));
      RTDF_TRY(layernorm_rows_f32(s, static_cast<const float*>(nxt), (long long)B * Lout, 512, f.ln.g, f.ln.b, 1e-5f,
                                  ACT_GELU, static_cast<float*>(nxt), nullptr));
    }
    void* t = cur;
    cur = nxt;
    nxt = t;
  }
  // LayerNorm(512) -> post_extract_proj -> x (fp32 stream) [+ bf16 copy for the pos-conv]
  if (bf) {
    RTDF_TRY(layernorm_rows_bf16(s, static_cast<const bf16*>(cur), M, 512, c->fp_ln.g, c->fp_ln.b, 1e-5f, ACT_NONE,
                                 nullptr, static_cast<bf16*>(w.featln)));
  } else {
    RTDF_TRY(layernorm_rows_f32(s, static_cast<const float*>(cur), M, 512, c->fp_ln.g, c->fp_ln.b, 1e-5f, ACT_NONE,
                                static_cast<float*>(w.featln), nullptr));
  }
  {
    TcEpilogue e;
    e.bias = c->proj.b;
    e.out_f32 = w.x;
    e.ld_f32 = 1024;
    if (bf) {
      e.out_bf16 = static_cast<bf16*>(w.xb);
      e.ld_bf16 = 1024;
    }
    RTDF_TRY(linear(c, s, w.featln, M, c->proj, e));
  }
  // x += GELU(pos_conv(x))
  if (bf && posconv_slab_enabled()) {
    RTDF_TRY(posconv_tc(s, w.x, static_cast<const bf16*>(w.xb), B, T, c->pos.wb, c->pos.b));
  } else if (bf) {
    TcOperandA a;
    a.ptr = static_cast<const bf16*>(w.xb);
    a.k_extent = 1024;
    a.rows_per_batch = T;
    a.batches = B;
    a.row_stride = 1024;
    a.batch_stride = (long long)T * 1024;
    TcEpilogue e;
    e.bias = c->pos.b;
    e.act = ACT_GELU;
    e.resid = w.x;
    e.ldr = 1024;
    e.out_f32 = w.x;
    e.ld_f32 = 1024;
    RTDF_TRY(tc_gemm(s, a, c->pos.wb, 1024, 8192, TC_POSCONV, 64, e));
  } else {
    RTDF_CHECK_CUDA(cudaMemcpyAsync(w.xb, w.x, (size_t)M * 1024 * 4, cudaMemcpyDeviceToDevice, s));
    RTDF_TRY(posconv_f32(s, w.x, static_cast<const float*>(w.xb), B, T, c->pos.w, c->pos.b));
  }
  // transformer layers (pre-LN).  A LayerNorm runs as its own row kernel (at the HBM roofline: 13 us for 12,736 rows).
  // Streaming chunks: out_proj / fc2 (16 output tiles per 128 rows, K up to 4096) split K over the idle SMs; the partial
  // sums land in w.partials and the LayerNorm that follows adds them to x in split order (deterministic).
  // RTDF_LN_FOLD=1 (bf16, opt-in, measured slower -- profiles/r02_ln_fold_experiment.txt): no LayerNorm kernels at all;
  // LN(x) W^T + b is evaluated by the projection that follows as rstd_i (bf16(x) W'^T - mean_i c) + d (weights folded at
  // pack time, gemm_tc.cuh TcEpilogue::fold_*), and the bf16 copy of the residual stream plus the per-row (sum, sum of
  // squares) come out of the epilogue of the residual GEMM that produced x.
  // Zig-zag traversal (bf16, large batches): every kernel of the layer stack walks its row blocks in the opposite direction
  // of its predecessor, so it starts on the rows that were written last and are still in L2 (each kernel's working set,
  // 50 - 130 MB, is of the order of the L2 itself).  Tile order does not touch the arithmetic: results are bit-identical.
  const bool zigzag = bf && zigzag_enabled() && !skinny_rows(c, M);
  bool rev = false;
  auto next_dir = [&]() { rev = zigzag ? !rev : false; return rev; };
  const bool fold = bf && c->ln_fold;
  const bool splitk = bf && !layer_taps && skinny_rows(c, M) && w.partials;
  const int sp_out = splitk ? tc_plan_splits(M, 1024, 1024) : 1, sp_fc2 = splitk ? tc_plan_splits(M, 1024, 4096) : 1;
  int pending = 0;      // K-split partials of the previous residual GEMM not yet folded into x
  auto layer_ln = [&](const Norm& n, float* of32, bf16* ob16) -> int {
    if (pending > 1) {
      const int np = pending;
      pending = 0;
      return layernorm_accum_rows(s, w.x, w.partials, np, M, n.g, n.b, 1e-5f, of32, ob16);
    }
    return layernorm_rows_f32(s, w.x, M, 1024, n.g, n.b, 1e-5f, ACT_NONE, of32, ob16, next_dir());
  };
  auto fold_prep = [&]() -> int {          // folded mode: xb / stats for the next projection, if not produced yet
    if (pending > 1) {
      const int np = pending;
      pending = 0;
      return cast_stats_rows(s, w.x, w.partials, np, M, static_cast<bf16*>(w.xb), w.stats);
    }
    return RTDF_OK;
  };
  // input of a projection that sits behind a LayerNorm: normalised rows in w.xb, or (folded) the epilogue terms
  auto ln_input = [&](const Norm& ln, const Lin& L, const float* cvec, const float* dvec, TcEpilogue& e) -> int {
    if (fold) {
      RTDF_TRY(fold_prep());
      e.bias = dvec;
      e.fold_c = cvec;
      e.fold_stats = w.stats;
      return RTDF_OK;
    }
    e.bias = L.b;
    return layer_ln(ln, bf ? nullptr : static_cast<float*>(w.xb), bf ? static_cast<bf16*>(w.xb) : nullptr);
  };
  auto residual_gemm = [&](const void* A, const Lin& L, int splits) -> int {   // x += A W^T + b
    TcEpilogue e;
    e.bias = L.b;
    if (splits > 1) {
      e.partials = w.partials;
      pending = splits;
      return tc_gemm(s, plainA(A, M, L.k), L.wb, L.n, L.k, TC_PLAIN, 64, e);
    }
    e.resid = w.x;
    e.ldr = 1024;
    e.out_f32 = w.x;
    e.ld_f32 = 1024;
    e.reverse_tiles = next_dir();
    e.a_cache_hint = bf ? a_hint_mode() : 0;
    if (fold) {
      e.xb_out = static_cast<bf16*>(w.xb);
      e.stats_out = w.stats;
      return linear(c, s, A, M, L, e, /*wide_tiles=*/true);
    }
    return linear(c, s, A, M, L, e);
  };
  const size_t n_layers = c->layers.size();
  if (c->layer_stack && c->stack_layers && !layer_taps && skinny_rows(c, M) && M <= kStackMaxRows && w.partials) {
    // <= 64 frames in flight (batch 1 x 1 s): all layers + the final LayerNorm in one persistent kernel
    StackParams sp;
    sp.layers = c->stack_layers;
    sp.n_layers = (int)n_layers;
    sp.R = (int)M;
    sp.B = B;
    sp.T = T;
    sp.x = w.x;
    sp.xn = static_cast<bf16*>(w.xb);
    sp.qkv = static_cast<bf16*>(w.qkv);
    sp.att = static_cast<bf16*>(w.attn);
    sp.h = static_cast<bf16*>(w.hbuf);
    sp.part = w.partials;
    sp.feats = feats;
    sp.gF = c->enc_ln.g;
    sp.bF = c->enc_ln.b;
    sp.wmaps = c->stack_wmaps;
    sp.big_boxes = c->stack_big_boxes;
    sp.impl = c->stack_impl;
    sp.sync = c->stack_sync;
    sp.fault = c->stack_fault;
    return layer_stack_bf16(s, sp);
  }
  const size_t tap_bytes = (size_t)M * 1024 * sizeof(float);
  if (layer_taps) RTDF_CHECK_CUDA(cudaMemcpyAsync(layer_taps, w.x, tap_bytes, cudaMemcpyDeviceToDevice, s));
  if (fold) RTDF_TRY(cast_stats_rows(s, w.x, nullptr, 0, M, static_cast<bf16*>(w.xb), w.stats));   // x after the pos-conv
  for (size_t l = 0; l < n_layers; ++l) {
    const XlsrLayer& L = c->layers[l];
    {
      TcEpilogue e;
      RTDF_TRY(ln_input(L.ln1, L.qkv, L.qkv_c, L.qkv_d, e));
      if (bf) { e.out_bf16 = static_cast<bf16*>(w.qkv); e.ld_bf16 = 3072; }
      else { e.out_f32 = static_cast<float*>(w.qkv); e.ld_f32 = 3072; }
      e.reverse_tiles = next_dir();
      e.a_cache_hint = bf ? a_hint_mode() : 0;
      RTDF_TRY(linear(c, s, w.xb, M, L.qkv, e));
    }
    if (bf) {
      if (T > 512)   // beyond the tcgen05 kernel's 512 key columns of tensor memory (10.2 s of audio): SIMT kernel
        RTDF_TRY(attention_simt_bf16(s, static_cast<const bf16*>(w.qkv), static_cast<bf16*>(w.attn), B, T, 16));
      else if (c->d.attention_impl == 0)
        RTDF_TRY(attention_ws(s, static_cast<const bf16*>(w.qkv), static_cast<bf16*>(w.attn), B, T, 16, next_dir()));
      else if (c->d.attention_impl == 2)
        RTDF_TRY(attention_tc(s, static_cast<const bf16*>(w.qkv), static_cast<bf16*>(w.attn), B, T, 16));
      else
        RTDF_TRY(attention_simt_bf16(s, static_cast<const bf16*>(w.qkv), static_cast<bf16*>(w.attn), B, T, 16));
    } else {
      RTDF_TRY(attention_simt_f32(s, static_cast<const float*>(w.qkv), static_cast<float*>(w.attn), B, T, 16));
    }
    RTDF_TRY(residual_gemm(w.attn, L.out, sp_out));
    {
      TcEpilogue e;
      e.act = ACT_GELU;
      RTDF_TRY(ln_input(L.ln2, L.fc1, L.fc1_c, L.fc1_d, e));
      if (bf) { e.out_bf16 = static_cast<bf16*>(w.hbuf); e.ld_bf16 = 4096; }
      else { e.out_f32 = static_cast<float*>(w.hbuf); e.ld_f32 = 4096; }
      e.reverse_tiles = next_dir();
      e.a_cache_hint = bf ? a_hint_mode() : 0;
      RTDF_TRY(linear(c, s, w.xb, M, L.fc1, e));
    }
    RTDF_TRY(residual_gemm(w.hbuf, L.fc2, sp_fc2));
    if (layer_taps) {  // output of encoder.layers[l] (the KD hook point, trainer.py:176-195)
      if (fold) RTDF_TRY(fold_prep());
      RTDF_CHECK_CUDA(cudaMemcpyAsync(layer_taps + (l + 1) * (size_t)M * 1024, w.x, tap_bytes, cudaMemcpyDeviceToDevice, s));
    }
  }
  RTDF_TRY(layer_ln(c->enc_ln, feats, nullptr));
  return RTDF_OK;
}

// ---------------------------------------------------------------------------------------------
// AASIST back-end
// ---------------------------------------------------------------------------------------------
static GraphView view(const float* p, int n, long long bs) {
  GraphView g;
  g.ptr = p;
  g.n = n;
  g.batch_stride = bs;
  return g;
}

static int run_aasist_encoder_simt(rtdf_ctx* c, cudaStream_t s, int B, int Tp, const AasistWs& w) {
  const AasistW& a = c->aasist;
  int cur = 0;
  for (int i = 0; i < 6; ++i) {
    const ResBlockW& b = a.blocks[i];
    const int tmp = (cur + 1) & 3, idt = (cur + 2) & 3, out = (cur + 3) & 3;
    Conv2dArgs c1;
    c1.in = w.r[cur]; c1.Ci = b.ci; c1.H = 42; c1.W = Tp; c1.w = b.conv1_w; c1.bias = b.conv1_b; c1.Co = b.co;
    c1.KH = 2; c1.pad_h = 1; c1.s1 = b.bn2.g; c1.t1 = b.bn2.b; c1.act1 = ACT_SELU; c1.out = w.r[tmp];
    RTDF_TRY(aasist_conv2d(s, c1, B));
    const float* resid = w.r[cur];
    if (b.ds_w) {
      Conv2dArgs cd;
      cd.in = w.r[cur]; cd.Ci = b.ci; cd.H = 42; cd.W = Tp; cd.w = b.ds_w; cd.bias = b.ds_b; cd.Co = b.co;
      cd.KH = 1; cd.pad_h = 0; cd.out = w.r[idt];
      RTDF_TRY(aasist_conv2d(s, cd, B));
      resid = w.r[idt];
    }
    Conv2dArgs c2;
    c2.in = w.r[tmp]; c2.Ci = b.co; c2.H = 43; c2.W = Tp; c2.w = b.conv2_w; c2.bias = b.conv2_b; c2.Co = b.co;
    c2.KH = 2; c2.pad_h = 0; c2.resid = resid; c2.out = w.r[out];
    if (i == 5) { c2.s2 = a.first_bn1.g; c2.t2 = a.first_bn1.b; c2.act2 = ACT_SELU; }  // xlsr_aasist.py:100-101
    RTDF_TRY(aasist_conv2d(s, c2, B));
    cur = out;
  }
  const float* x = w.r[cur];
  RTDF_TRY(aasist_attn_map(s, x, B, 42, Tp, a.att_w1t, a.att_b1, a.att_bn.g, a.att_bn.b, a.att_w2t, a.att_b2, w.wmap));
  RTDF_TRY(aasist_attn_pool(s, x, w.wmap, B, 42, Tp, a.pos_S, w.eS, w.eT));
  return RTDF_OK;
}

// Residual encoder + attention map + attention pooling on zero-padded channels-last planes with the shifted-row
// tcgen05 conv (conv_tc.cuh).  Plane conventions: 42-row tensors sit at hp = h + 1, conv1's 43-row output at hp = h.
static int run_aasist_encoder_tc(rtdf_ctx* c, cudaStream_t s, int B, int Tp, const AasistWs& w) {
  const AasistW& a = c->aasist;
  const int Hp = kAasistHp, Wp = Tp + 2;
  const long long rows = (long long)B * Hp * Wp;
  const int nsplit = c->d.aasist_conv_impl == 1 ? 1 : 3;
  auto base = [&](const TcConvW& wt) {
    ConvTcArgs g;
    g.rows = rows; g.Hp = Hp; g.Wp = Wp;
    g.w_hi = wt.hi; g.w_lo = wt.lo; g.ci = wt.ci; g.co = wt.co; g.n_chunks = wt.n_chunks;
    return g;
  };
  // conv1: (2,3) pad (1,1), 43 output rows at hp = h: input row hp + kh.  conv2: (2,3) pad (0,1), output at hp = h + 1:
  // input row hp - 1 + kh.  downsample: (1,3) pad (0,1), same row.
  auto taps_conv1 = [&](ConvTcArgs& g) { for (int kh = 0; kh < 2; ++kh) for (int kw = 0; kw < 3; ++kw) g.shift[kh * 3 + kw] = kh * Wp + kw - 1; g.hp_lo = 0; g.hp_hi = 42; };
  auto taps_conv2 = [&](ConvTcArgs& g) { for (int kh = 0; kh < 2; ++kh) for (int kw = 0; kw < 3; ++kw) g.shift[kh * 3 + kw] = (kh - 1) * Wp + kw - 1; g.hp_lo = 1; g.hp_hi = 42; };
  auto taps_ds = [&](ConvTcArgs& g) { for (int kw = 0; kw < 3; ++kw) g.shift[kw] = kw - 1; g.hp_lo = 1; g.hp_hi = 42; };
  int cur = 0;
  for (int i = 0; i < 6; ++i) {
    const ResBlockW& b = a.blocks[i];
    const AasistWs::Plane& X = w.pl[cur];
    const AasistWs::Plane& O = w.pl[cur ^ 1];
    const float* resid = X.f;
    if (i == 0) {
      RTDF_TRY(conv_tc_block0(s, w.r[0], B, Tp, Hp, Wp, b.conv1_raw, b.conv1_b, b.bn2.g, b.bn2.b, b.ds_raw, b.ds_b,
                              w.y_hi, w.y_lo, w.idt));
      resid = w.idt;
    } else {
      ConvTcArgs c1 = base(b.tc1);
      taps_conv1(c1);
      c1.in_hi = X.hi; c1.in_lo = X.lo;
      c1.bias = b.conv1_b; c1.s1 = b.bn2.g; c1.t1 = b.bn2.b; c1.act1 = ACT_SELU;
      c1.out_hi = w.y_hi; c1.out_lo = w.y_lo;
      RTDF_TRY(conv_tc(s, c1, nsplit));
      if (b.tcd.hi) {
        ConvTcArgs cd = base(b.tcd);
        taps_ds(cd);
        cd.in_hi = X.hi; cd.in_lo = X.lo;
        cd.bias = b.ds_b;
        cd.out_f32 = w.idt;
        RTDF_TRY(conv_tc(s, cd, nsplit));
        resid = w.idt;
      }
    }
    ConvTcArgs c2 = base(b.tc2);
    taps_conv2(c2);
    c2.in_hi = w.y_hi; c2.in_lo = w.y_lo;
    c2.bias = b.conv2_b; c2.resid = resid;
    if (i == 5) { c2.s2 = a.first_bn1.g; c2.t2 = a.first_bn1.b; c2.act2 = ACT_SELU; }   // xlsr_aasist.py:100-101
    c2.out_f32 = O.f; c2.out_hi = O.hi; c2.out_lo = O.lo;
    RTDF_TRY(conv_tc(s, c2, nsplit));
    cur ^= 1;
  }
  const AasistWs::Plane& X = w.pl[cur];
  {
    ConvTcArgs g1 = base(a.att1);
    g1.hp_lo = 1; g1.hp_hi = 42;
    g1.in_hi = X.hi; g1.in_lo = X.lo;
    g1.bias = a.att_b1; g1.act1 = ACT_SELU;
    g1.out_hi = w.hid_hi; g1.out_lo = w.hid_lo;
    RTDF_TRY(conv_tc(s, g1, nsplit));
    ConvTcArgs g2 = base(a.att2);
    g2.hp_lo = 1; g2.hp_hi = 42;
    g2.sub[0] = 0; g2.sub[1] = 1;
    g2.in_hi = w.hid_hi; g2.in_lo = w.hid_lo;
    g2.bias = a.att_b2_folded;
    g2.out_f32 = w.wmap;
    RTDF_TRY(conv_tc(s, g2, nsplit));
  }
  return attn_pool_planes(s, X.f, w.wmap, B, 42, Tp, Hp, Wp, a.pos_S, w.eS, w.eT);
}

static int gat_rows(const rtdf_ctx* c, cudaStream_t s, int D, int DO, const GraphView& x, int B, int n1,
                    const GatRowWeights& w, float* out, long long out_bs, const float* master_in, long long master_stride,
                    const GatRowWeights* wM, float* master_out) {
  if (aasist_tc(c) && c->d.gat_impl == 0)
    return aasist_gat_rows_mma(s, D, DO, x, B, n1, w, out, out_bs, master_in, master_stride, wM, master_out);
  return aasist_gat_rows(s, D, DO, x, B, n1, w, out, out_bs, master_in, master_stride, wM, master_out);
}

static int run_aasist(rtdf_ctx* c, cudaStream_t s, const float* feats, int B, int T, const AasistWs& w, float* logits,
                      const rtdf_taps* taps) {
  const AasistW& a = c->aasist;
  const bool bf = c->d.precision == RTDF_PREC_BF16;
  const long long M = (long long)B * T;
  const int Tp = T / 3, kT = Tp / 2 > 0 ? Tp / 2 : 1, kT2 = kT / 2 > 0 ? kT / 2 : 1;
  // T' = T/3 temporal nodes (graph kernels: up to 256 nodes; shifted-row convs: slab of 128 + T' + 4 plane rows through one
  // or two TMA boxes).  bf16: up to 512 frames (10.2 s), the reach of the tcgen05 attention; fp32 verification mode: 386.
  const int max_tp = aasist_tc(c) ? 170 : 128;
  RTDF_REQUIRE(Tp >= 1 && Tp <= max_tp, "AASIST back-end supports 3..%d frames (%.1f s of audio) in this mode, got T = %d",
               max_tp * 3 + 2, (max_tp * 3 + 2) * 0.02, T);
  {
    TcEpilogue e;
    e.bias = a.LL.b;
    e.out_f32 = w.z;
    e.ld_f32 = 128;
    const void* A = feats;
    if (bf) {
      RTDF_TRY(cast_f32_to_bf16(s, feats, static_cast<bf16*>(w.featsb), M * 1024));
      A = w.featsb;
    }
    RTDF_TRY(linear(c, s, A, M, a.LL, e));
  }
  RTDF_TRY(aasist_stem(s, w.z, B, T, a.first_bn_s, a.first_bn_t, w.r[0]));
  if (aasist_tc(c)) {
    RTDF_TRY(run_aasist_encoder_tc(c, s, B, Tp, w));
  } else {
    RTDF_TRY(run_aasist_encoder_simt(c, s, B, Tp, w));
  }
  // The spectral and the temporal graph, and later the two heterogeneous branches, are independent until the read-out:
  // they run on two streams (fork / join through events -- capturable, ordered with respect to the caller's stream), so
  // their small latency-bound kernels overlap instead of queueing behind each other.
  cudaStream_t s2 = c->side_stream ? c->side_stream : s;
  auto fork = [&]() -> int {
    if (s2 == s) return RTDF_OK;
    RTDF_CHECK_CUDA(cudaEventRecord(c->ev_fork, s));
    RTDF_CHECK_CUDA(cudaStreamWaitEvent(s2, c->ev_fork, 0));
    return RTDF_OK;
  };
  auto join = [&]() -> int {
    if (s2 == s) return RTDF_OK;
    RTDF_CHECK_CUDA(cudaEventRecord(c->ev_join, s2));
    RTDF_CHECK_CUDA(cudaStreamWaitEvent(s, c->ev_join, 0));
    return RTDF_OK;
  };
  RTDF_TRY(fork());
  RTDF_TRY(gat_rows(c, s, 64, 64, view(w.eS, 42, 42 * 64), B, 42, a.gat_S, w.gS, 42 * 64, nullptr, 0, nullptr, nullptr));
  RTDF_TRY(aasist_graph_pool(s, 64, view(w.gS, 42, 42 * 64), B, a.pool_S.w, a.pool_S.b, 21, w.oS, w.idxS));
  RTDF_TRY(gat_rows(c, s2, 64, 64, view(w.eT, Tp, (long long)Tp * 64), B, Tp, a.gat_T, w.gT, (long long)Tp * 64, nullptr, 0, nullptr, nullptr));
  RTDF_TRY(aasist_graph_pool(s2, 64, view(w.gT, Tp, (long long)Tp * 64), B, a.pool_T.w, a.pool_T.b, kT, w.oT, w.idxT));
  RTDF_TRY(join());
  const HsGalW* l1[2] = {&a.st11, &a.st21};
  const HsGalW* l2[2] = {&a.st12, &a.st22};
  const PoolW* pS[2] = {&a.pool_hS1, &a.pool_hS2};
  const PoolW* pT[2] = {&a.pool_hT1, &a.pool_hT2};
  const float* master[2] = {a.master1, a.master2};
  const int n = kT + 21, n2 = kT2 + 10;
  RTDF_TRY(fork());
  for (int i = 0; i < 2; ++i) {
    const AasistWs::Br& r = w.br[i];
    cudaStream_t sb = i == 0 ? s : s2;
    RTDF_TRY(aasist_type_proj(sb, 64, view(w.oT, kT, (long long)kT * 64), view(w.oS, 21, 21 * 64), B, l1[i]->t1_wt,
                              l1[i]->t1_b, l1[i]->t2_wt, l1[i]->t2_b, r.hx));
    RTDF_TRY(gat_rows(c, sb, 64, 32, view(r.hx, n, (long long)n * 64), B, kT, l1[i]->rows, r.hy, (long long)n * 32,
                             master[i], 0, &l1[i]->master, r.ma));
    RTDF_TRY(aasist_graph_pool(sb, 32, view(r.hy + (long long)kT * 32, 21, (long long)n * 32), B, pS[i]->w, pS[i]->b, 10, r.pS, nullptr));
    RTDF_TRY(aasist_graph_pool(sb, 32, view(r.hy, kT, (long long)n * 32), B, pT[i]->w, pT[i]->b, kT2, r.pT, nullptr));
    RTDF_TRY(aasist_type_proj(sb, 32, view(r.pT, kT2, (long long)kT2 * 32), view(r.pS, 10, 10 * 32), B, l2[i]->t1_wt,
                              l2[i]->t1_b, l2[i]->t2_wt, l2[i]->t2_b, r.hx2));
    RTDF_TRY(gat_rows(c, sb, 32, 32, view(r.hx2, n2, (long long)n2 * 32), B, kT2, l2[i]->rows, r.hy2, (long long)n2 * 32,
                             r.ma, 32, &l2[i]->master, r.mb));
  }
  RTDF_TRY(join());
  ReadoutArgs ro;
  ro.T1 = view(w.br[0].pT, kT2, (long long)kT2 * 32);
  ro.Ta1 = view(w.br[0].hy2, kT2, (long long)n2 * 32);
  ro.S1 = view(w.br[0].pS, 10, 10 * 32);
  ro.T2 = view(w.br[1].pT, kT2, (long long)kT2 * 32);
  ro.Ta2 = view(w.br[1].hy2, kT2, (long long)n2 * 32);
  ro.S2 = view(w.br[1].pS, 10, 10 * 32);
  ro.Sa2 = view(w.br[1].hy2 + (long long)kT2 * 32, 10, (long long)n2 * 32);
  ro.m1a = w.br[0].ma; ro.m1b = w.br[0].mb; ro.m2a = w.br[1].ma; ro.m2b = w.br[1].mb;
  ro.w = a.out_w; ro.b = a.out_b; ro.logits = logits;
  ro.hidden = taps ? taps->hidden : nullptr;
  RTDF_TRY(aasist_readout(s, ro, B));
  if (taps && taps->idx_S) RTDF_CHECK_CUDA(cudaMemcpyAsync(taps->idx_S, w.idxS, (size_t)B * 21 * 4, cudaMemcpyDeviceToDevice, s));
  if (taps && taps->idx_T) RTDF_CHECK_CUDA(cudaMemcpyAsync(taps->idx_T, w.idxT, (size_t)B * kT * 4, cudaMemcpyDeviceToDevice, s));
  return RTDF_OK;
}

// ---------------------------------------------------------------------------------------------
// Conformer back-end (conformer_baseline.py:22-29, 54-64)
// ---------------------------------------------------------------------------------------------
static void plan_conformer(const rtdf_ctx* c, int B, int T, Bump& b, ConformerWs* w) {
  const size_t es = c->d.precision == RTDF_PREC_BF16 ? 2 : 4;
  const long long M = (long long)B * T, R = (long long)B * (T + 1);
  const int E = c->d.conf_emb;
  w->featsb = c->d.precision == RTDF_PREC_BF16 ? (void*)b.take<bf16>(M * 1024) : nullptr;
  w->z = b.take<float>(M * E);
  w->x = b.take<float>(R * E);
  w->yb = b.take<char>(R * E * es);
  w->h = b.take<char>(R * 4 * E * es);
  w->ab = b.take<char>(R * E * es);
  w->dw = b.take<char>(R * 2 * E * es);
}

static int conf_ln(const rtdf_ctx* c, cudaStream_t s, const float* x, long long rows, int E, const Norm& n, void* out) {
  const bool bf = c->d.precision == RTDF_PREC_BF16;
  return layernorm_rows_f32(s, x, rows, E, n.g, n.b, 1e-5f, ACT_NONE, bf ? nullptr : static_cast<float*>(out),
                            bf ? static_cast<bf16*>(out) : nullptr);
}

static void set_out(const rtdf_ctx* c, TcEpilogue& e, void* out, int ld) {
  if (c->d.precision == RTDF_PREC_BF16) {
    e.out_bf16 = static_cast<bf16*>(out);
    e.ld_bf16 = ld;
  } else {
    e.out_f32 = static_cast<float*>(out);
    e.ld_f32 = ld;
  }
}

static int conf_ff(const rtdf_ctx* c, cudaStream_t s, float* x, long long R, int E, const Norm& ln, const Lin& a,
                   const Lin& b, const ConformerWs& w) {
  RTDF_TRY(conf_ln(c, s, x, R, E, ln, w.yb));
  TcEpilogue e1;
  e1.bias = a.b;
  e1.act = ACT_SWISH;
  set_out(c, e1, w.h, 4 * E);
  RTDF_TRY(linear(c, s, w.yb, R, a, e1));
  TcEpilogue e2;   // x = x + 0.5 * (W h + b)
  e2.bias = b.b;
  e2.scale = 0.5f;
  e2.resid = x;
  e2.ldr = E;
  e2.out_f32 = x;
  e2.ld_f32 = E;
  return linear(c, s, w.h, R, b, e2);
}

static int run_conformer(rtdf_ctx* c, cudaStream_t s, const float* feats, int B, int T, const ConformerWs& w, float* logits) {
  const ConformerW& cf = c->conf;
  const bool bf = c->d.precision == RTDF_PREC_BF16;
  const int E = c->d.conf_emb, heads = c->d.conf_heads, dh = E / heads, K = c->d.conf_kernel, n = T + 1;
  const long long M = (long long)B * T, R = (long long)B * n;
  {
    TcEpilogue e;
    e.bias = cf.LL.b;
    e.out_f32 = w.z;
    e.ld_f32 = E;
    const void* A = feats;
    if (bf) {
      RTDF_TRY(cast_f32_to_bf16(s, feats, static_cast<bf16*>(w.featsb), M * 1024));
      A = w.featsb;
    }
    RTDF_TRY(linear(c, s, A, M, cf.LL, e));
  }
  RTDF_TRY(conformer_stem(s, w.z, cf.class_token, B, T, E, cf.first_bn_s, cf.first_bn_t, w.x));
  for (const ConformerBlockW& b : cf.blocks) {
    RTDF_TRY(conf_ff(c, s, w.x, R, E, b.ff1_ln, b.ff1_a, b.ff1_b, w));
    // attention
    RTDF_TRY(conf_ln(c, s, w.x, R, E, b.attn_ln, w.yb));
    {
      TcEpilogue e;
      set_out(c, e, w.h, 3 * E);
      RTDF_TRY(linear(c, s, w.yb, R, b.qkv, e));
    }
    if (bf)
      RTDF_TRY(conformer_attention_bf16(s, static_cast<const bf16*>(w.h), b.rel_pos, static_cast<bf16*>(w.ab), B, n, heads, dh));
    else
      RTDF_TRY(conformer_attention_f32(s, static_cast<const float*>(w.h), b.rel_pos, static_cast<float*>(w.ab), B, n, heads, dh));
    {
      TcEpilogue e;
      e.bias = b.attn_out.b;
      e.resid = w.x;
      e.ldr = E;
      e.out_f32 = w.x;
      e.ld_f32 = E;
      RTDF_TRY(linear(c, s, w.ab, R, b.attn_out, e));
    }
    // convolution module
    RTDF_TRY(conf_ln(c, s, w.x, R, E, b.conv_ln, w.yb));
    {
      TcEpilogue e;
      e.bias = b.pw1.b;
      set_out(c, e, w.h, 4 * E);
      RTDF_TRY(linear(c, s, w.yb, R, b.pw1, e));
    }
    if (bf)
      RTDF_TRY(conformer_glu_dwconv_bf16(s, static_cast<const bf16*>(w.h), static_cast<bf16*>(w.dw), B, n, 2 * E, K,
                                         b.dw_w, b.dw_b, b.dw_bn.g, b.dw_bn.b));
    else
      RTDF_TRY(conformer_glu_dwconv_f32(s, static_cast<const float*>(w.h), static_cast<float*>(w.dw), B, n, 2 * E, K,
                                        b.dw_w, b.dw_b, b.dw_bn.g, b.dw_bn.b));
    {
      TcEpilogue e;
      e.bias = b.pw2.b;
      e.resid = w.x;
      e.ldr = E;
      e.out_f32 = w.x;
      e.ld_f32 = E;
      RTDF_TRY(linear(c, s, w.dw, R, b.pw2, e));
    }
    RTDF_TRY(conf_ff(c, s, w.x, R, E, b.ff2_ln, b.ff2_a, b.ff2_b, w));
    RTDF_TRY(layernorm_rows_f32(s, w.x, R, E, b.post_ln.g, b.post_ln.b, 1e-5f, ACT_NONE, w.x, nullptr));
  }
  return conformer_head(s, w.x, B, n, E, cf.fc5.w, cf.fc5.b, logits);
}

static size_t backend_plan(const rtdf_ctx* c, int B, int T, Bump& b, AasistWs* aw, ConformerWs* cw) {
  if (c->d.backend == RTDF_BACKEND_AASIST) plan_aasist(c, B, T, b, aw);
  else if (c->d.backend == RTDF_BACKEND_CONFORMER) plan_conformer(c, B, T, b, cw);
  return b.off;
}

}  // namespace rtdf

// =================================================================================================
// C-ABI: lifecycle + forward
// =================================================================================================
extern "C" {

const char* rtdf_last_error(void) { return rtdf::get_error(); }

int rtdf_num_frames(int n) {
  Dims d;
  if (make_dims(1, n, &d) != RTDF_OK) return 0;
  return d.T;
}

int rtdf_create(rtdf_ctx** out, int device, const rtdf_model_desc* desc) {
  RTDF_REQUIRE(out && desc, "rtdf_create: null argument");
  RTDF_REQUIRE(desc->backend >= 0 && desc->backend <= 2, "rtdf_create: unknown backend %d", desc->backend);
  RTDF_REQUIRE(desc->n_layers >= 1 && desc->n_layers <= 24, "Number of layers must be at least 1 and at most 24.");
  RTDF_REQUIRE(desc->precision == RTDF_PREC_BF16 || desc->precision == RTDF_PREC_FP32, "rtdf_create: bad precision");
  int count = 0;
  RTDF_CHECK_CUDA(cudaGetDeviceCount(&count));
  RTDF_REQUIRE(device >= 0 && device < count, "rtdf_create: device %d not present (%d devices)", device, count);
  cudaDeviceProp prop;
  RTDF_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    rtdf::set_error("rtdf_create: device %d is sm_%d%d; this library contains sm_100a (B200) code only", device,
                    prop.major, prop.minor);
    return RTDF_ERR_UNSUPPORTED;
  }
  RTDF_CHECK_CUDA(cudaSetDevice(device));
  rtdf_ctx* c = new rtdf_ctx();
  c->device = device;
  c->d = *desc;
  if (c->d.conf_emb == 0) c->d.conf_emb = 144;
  if (c->d.conf_heads == 0) c->d.conf_heads = 4;
  if (c->d.conf_kernel == 0) c->d.conf_kernel = 31;
  if (c->d.conf_blocks == 0) c->d.conf_blocks = 4;
  {
    const char* e = getenv("RTDF_LN_FOLD");
    c->ln_fold = e && e[0] == '1' && c->d.precision == RTDF_PREC_BF16;
  }
  {  // RTDF_LAYER_STACK=0: streaming chunks of <= 64 frames keep the kernel-per-op chain instead of the persistent layer-stack kernel
    const char* e = getenv("RTDF_LAYER_STACK");
    c->layer_stack = !(e && e[0] == '0') && c->d.precision == RTDF_PREC_BF16 && !c->ln_fold;
    e = getenv("RTDF_STACK_IMPL");   // mma: its mma.sync variant (A/B timing)
    c->stack_impl = (e && e[0] == 'm') ? 1 : 0;
  }
  {  // second stream + fork / join events for the independent graph branches of the AASIST back-end (RTDF_BRANCH_STREAMS=0: off)
    const char* e = getenv("RTDF_BRANCH_STREAMS");
    if (!(e && e[0] == '0') && c->d.backend == RTDF_BACKEND_AASIST) {
      cudaError_t err = cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking);
      if (err == cudaSuccess) err = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
      if (err == cudaSuccess) err = cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
      if (err != cudaSuccess) {
        rtdf::set_error("rtdf_create: cannot create the branch stream / events: %s", cudaGetErrorString(err));
        rtdf_destroy(c);
        return RTDF_ERR_CUDA;
      }
    }
  }
  *out = c;
  return RTDF_OK;
}

int rtdf_load_weight(rtdf_ctx* c, const char* key, const void* data, const int64_t* shape, int ndim) {
  RTDF_REQUIRE(c && key && data && (shape || ndim == 0) && ndim >= 0 && ndim <= 8, "rtdf_load_weight: bad arguments");
  if (c->finalized) {
    rtdf::set_error("rtdf_load_weight: context already finalized");
    return RTDF_ERR_STATE;
  }
  RTDF_CHECK_CUDA(cudaSetDevice(c->device));
  Raw r;
  r.numel = 1;
  for (int i = 0; i < ndim; ++i) {
    RTDF_REQUIRE(shape[i] >= 0, "rtdf_load_weight: negative dimension");
    r.shape.push_back(shape[i]);
    r.numel *= shape[i];
  }
  std::string k(key);
  if (k.rfind("module.", 0) == 0) k = k.substr(7);  // DataParallel / DDP prefix (reference utils.py:13-43)
  auto it = c->raw.find(k);
  if (it != c->raw.end() && it->second.numel == r.numel) {
    r.p = it->second.p;
  } else {
    void* p = nullptr;
    RTDF_CHECK_CUDA(cudaMalloc(&p, (size_t)(r.numel > 0 ? r.numel : 1) * sizeof(float)));
    c->owned.push_back(p);
    r.p = static_cast<float*>(p);
  }
  RTDF_CHECK_CUDA(cudaMemcpy(r.p, data, (size_t)r.numel * sizeof(float), cudaMemcpyDefault));
  c->raw[k] = r;
  return RTDF_OK;
}

int rtdf_set_regime(rtdf_ctx* c, int regime) {
  RTDF_REQUIRE(c, "rtdf_set_regime: null context");
  RTDF_REQUIRE(regime == RTDF_REGIME_AUTO || regime == RTDF_REGIME_THROUGHPUT, "rtdf_set_regime: unknown regime %d", regime);
  c->regime = regime;
  return RTDF_OK;
}

int rtdf_finalize(rtdf_ctx* c) {
  RTDF_REQUIRE(c, "rtdf_finalize: null context");
  if (c->finalized) return RTDF_OK;
  RTDF_CHECK_CUDA(cudaSetDevice(c->device));
  RTDF_TRY(pack_xlsr(c));
  if (c->d.backend == RTDF_BACKEND_AASIST) RTDF_TRY(pack_aasist(c));
  if (c->d.backend == RTDF_BACKEND_CONFORMER) RTDF_TRY(pack_conformer(c));
  if (c->layer_stack) {
    std::vector<rtdf::StackLayer> host;
    for (const XlsrLayer& L : c->layers) {
      rtdf::StackLayer sl;
      sl.wqkv = L.qkv.wb; sl.wo = L.out.wb; sl.w1 = L.fc1.wb; sl.w2 = L.fc2.wb;
      sl.bqkv = L.qkv.b; sl.bo = L.out.b; sl.b1 = L.fc1.b; sl.b2 = L.fc2.b;
      sl.g1 = L.ln1.g; sl.be1 = L.ln1.b; sl.g2 = L.ln2.g; sl.be2 = L.ln2.b;
      host.push_back(sl);
    }
    void* dev = nullptr;
    RTDF_CHECK_CUDA(cudaMalloc(&dev, host.size() * sizeof(rtdf::StackLayer) + 256));
    c->owned.push_back(dev);
    RTDF_CHECK_CUDA(cudaMemcpy(dev, host.data(), host.size() * sizeof(rtdf::StackLayer), cudaMemcpyHostToDevice));
    c->stack_layers = static_cast<rtdf::StackLayer*>(dev);
    {
      std::vector<CUtensorMap> maps(4 * host.size());
      RTDF_TRY(rtdf::layer_stack_build_wmaps(host.data(), (int)host.size(), maps.data(), &c->stack_big_boxes));
      void* dm = nullptr;
      RTDF_CHECK_CUDA(cudaMalloc(&dm, maps.size() * sizeof(CUtensorMap)));
      c->owned.push_back(dm);
      RTDF_CHECK_CUDA(cudaMemcpy(dm, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
      c->stack_wmaps = static_cast<CUtensorMap*>(dm);
    }
    void* words = nullptr;
    RTDF_CHECK_CUDA(cudaMalloc(&words, 256));
    c->owned.push_back(words);
    RTDF_CHECK_CUDA(cudaMemset(words, 0, 256));
    c->stack_sync = static_cast<unsigned*>(words);
    void* flag = nullptr;
    RTDF_CHECK_CUDA(cudaHostAlloc(&flag, sizeof(int), cudaHostAllocMapped));
    c->stack_fault = static_cast<int*>(flag);
    *c->stack_fault = 0;
  }
  RTDF_CHECK_CUDA(cudaDeviceSynchronize());
  if (!c->scratch.empty()) {   // bf16 mode: release the fp32 sources of the packed GEMM weights
    std::set<void*> dead(c->scratch.begin(), c->scratch.end());
    for (auto it = c->raw.begin(); it != c->raw.end();)
      it = dead.count(it->second.p) ? c->raw.erase(it) : std::next(it);
    std::vector<void*> keep;
    for (void* p : c->owned) {
      if (dead.count(p)) cudaFree(p);
      else keep.push_back(p);
    }
    c->owned.swap(keep);
    c->scratch.clear();
  }
  c->finalized = true;
  return RTDF_OK;
}

void rtdf_destroy(rtdf_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  for (void* p : c->owned) cudaFree(p);
  if (c->stack_fault) cudaFreeHost(c->stack_fault);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->side_stream) cudaStreamDestroy(c->side_stream);
  delete c;
}

int rtdf_workspace_bytes(const rtdf_ctx* c, int B, int N, size_t* out) {
  RTDF_REQUIRE(c && out, "rtdf_workspace_bytes: null argument");
  Dims d;
  RTDF_TRY(make_dims(B, N, &d));
  Bump b;
  FrontWs fw;
  AasistWs aw;
  ConformerWs cw;
  plan_front(c, d, true, true, b, &fw);
  backend_plan(c, B, d.T, b, &aw, &cw);
  *out = b.off + 256;
  return RTDF_OK;
}

static int check_ready(const rtdf_ctx* c) {
  RTDF_REQUIRE(c, "null context");
  if (!c->finalized) {
    rtdf::set_error("context not finalized: call rtdf_finalize() after loading the weights");
    return RTDF_ERR_STATE;
  }
  if (c->stack_fault && *static_cast<volatile int*>(c->stack_fault)) {
    rtdf::set_error("an earlier forward failed: the layer-stack kernel timed out at a grid barrier (its 128 CTAs were not "
                    "co-resident); results since then are invalid.  RTDF_LAYER_STACK=0 selects the kernel-per-op chain");
    return RTDF_ERR_STATE;
  }
  return RTDF_OK;
}

static int run_backend(rtdf_ctx* c, cudaStream_t s, const float* feats, int B, int T, float* logits,
                       const rtdf_taps* taps, const AasistWs& aw, const ConformerWs& cw) {
  if (c->d.backend == RTDF_BACKEND_AASIST) return run_aasist(c, s, feats, B, T, aw, logits, taps);
  if (c->d.backend == RTDF_BACKEND_CONFORMER) return run_conformer(c, s, feats, B, T, cw, logits);
  rtdf::set_error("context was created without a back-end");
  return RTDF_ERR_STATE;
}

int rtdf_forward(rtdf_ctx* c, const float* wav, int B, int N, int preemph_on, float coef, float* logits, void* ws,
                 size_t ws_bytes, const rtdf_taps* taps, void* stream) {
  RTDF_TRY(check_ready(c));
  RTDF_REQUIRE(wav && logits && ws, "rtdf_forward: null buffer");
  Dims d;
  RTDF_TRY(make_dims(B, N, &d));
  Bump b;
  b.base = static_cast<char*>(ws);
  FrontWs fw;
  AasistWs aw{};
  ConformerWs cw{};
  plan_front(c, d, preemph_on != 0, true, b, &fw);
  backend_plan(c, B, d.T, b, &aw, &cw);
  RTDF_REQUIRE(b.off <= ws_bytes, "rtdf_forward: workspace too small (%zu < %zu bytes)", ws_bytes, b.off);
  RTDF_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "rtdf_forward: workspace must be 256-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  pdl_set_auto(d.M <= kPdlRows);
  RTDF_TRY(run_frontend(c, s, wav, d, preemph_on, coef, fw, fw.feats, taps ? taps->layers : nullptr));
  if (taps && taps->feats)
    RTDF_CHECK_CUDA(cudaMemcpyAsync(taps->feats, fw.feats, (size_t)d.M * 1024 * 4, cudaMemcpyDeviceToDevice, s));
  return run_backend(c, s, fw.feats, B, d.T, logits, taps, aw, cw);
}

int rtdf_frontend(rtdf_ctx* c, const float* wav, int B, int N, int preemph_on, float coef, float* feats, void* ws,
                  size_t ws_bytes, void* stream) {
  RTDF_TRY(check_ready(c));
  RTDF_REQUIRE(wav && feats && ws, "rtdf_frontend: null buffer");
  Dims d;
  RTDF_TRY(make_dims(B, N, &d));
  Bump b;
  b.base = static_cast<char*>(ws);
  FrontWs fw;
  plan_front(c, d, preemph_on != 0, false, b, &fw);
  RTDF_REQUIRE(b.off <= ws_bytes, "rtdf_frontend: workspace too small (%zu < %zu bytes)", ws_bytes, b.off);
  RTDF_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "rtdf_frontend: workspace must be 256-byte aligned");
  pdl_set_auto(d.M <= kPdlRows);
  return run_frontend(c, static_cast<cudaStream_t>(stream), wav, d, preemph_on, coef, fw, feats);
}

int rtdf_backend(rtdf_ctx* c, const float* feats, int B, int T, float* logits, void* ws, size_t ws_bytes,
                 const rtdf_taps* taps, void* stream) {
  RTDF_TRY(check_ready(c));
  RTDF_REQUIRE(feats && logits && ws && B >= 1 && T >= 1, "rtdf_backend: bad arguments");
  Bump b;
  b.base = static_cast<char*>(ws);
  AasistWs aw{};
  ConformerWs cw{};
  backend_plan(c, B, T, b, &aw, &cw);
  RTDF_REQUIRE(b.off <= ws_bytes, "rtdf_backend: workspace too small (%zu < %zu bytes)", ws_bytes, b.off);
  RTDF_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "rtdf_backend: workspace must be 256-byte aligned");
  pdl_set_auto((long long)B * T <= kPdlRows);
  return run_backend(c, static_cast<cudaStream_t>(stream), feats, B, T, logits, taps, aw, cw);
}

}  // extern "C"
