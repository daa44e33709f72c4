// Training path, last step: reduce the split-K slabs of the weight-gradient kernel (deterministic order) and apply the
// chain rule through the two folded input layers, producing gradients in the reference's parameter layouts
// (PyTorch [out, in]) for every tensor of the MLP + the gradient of the per-frame latent (which PyTorch autograd, or
// s2l's AudioNet backward, carries on into encoder_conv / encoder_fc1).  See s2l_train.cuh for the algebra.
// Also the C ABI of the training path (s2l_train_*).
#include "s2l_train.cuh"

namespace s2l {

int launch_mlp_tc_train(const void* blob, const PointSrc& src, int n_frames, const float* frame_bias, float* rgb,
                        __nv_bfloat16* save_h, __nv_bfloat16* save_pe, uint32_t* save_mask, cudaStream_t st, float* raw_out = nullptr);
int launch_dgrad_tc(const void* blob, const PointSrc& src, int n_frames, const float* d_rgb, const TrainBufs& B, cudaStream_t st,
                    const float* d_out_rows = nullptr);
int launch_wgrad_tc(const WgPlan& plan, const TrainBufs& B, cudaStream_t st);

struct GradPtrs {
  float* g[S2L_NUM_PARAMS];     // device pointers in S2L_P_* order; AudioNet entries unused
  float* d_latent;              // [F, 64]
};

// small reduced quantities kept in the workspace between the two finalize kernels (float offsets)
constexpr int R_M0 = 0;                        // [256][64]
constexpr int R_M5 = R_M0 + 256 * 64;
constexpr int R_S0 = R_M5 + 256 * 64;          // [F][256] per-frame column sums of dPre0, then S5, g0, g5
__host__ __device__ inline long long red_floats(int F) { return R_S0 + 4ll * F * 256; }

// ---- kernel 1: slab reduction.  One thread per output element.
__global__ void __launch_bounds__(256) wg_reduce_kernel(WgPlan pl, const float* __restrict__ partials, GradPtrs G, float* __restrict__ red,
                                                        const float* __restrict__ dbout_part, int dbout_rows, int out_ch) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nL = 7ll * 65536, nLb = 7ll * 256, nM = 2ll * 16384, nS = 2ll * pl.F * 256, nO = 256ll * 4;
  if (gid < nL) {
    const int job = (int)(gid >> 16), e = (int)(gid & 65535);
    float s = 0.f;
    for (int sl = 0; sl < pl.SL; ++sl) s += partials[pl.off_L(job, sl) + e];
    const int l = job + 1;
    if (l == 5) G.g[S2L_P_PTS0_W + 10][(e >> 8) * 512 + 256 + (e & 255)] = s;     // pts_linears.5.weight[:, 256:]
    else G.g[S2L_P_PTS0_W + 2 * l][e] = s;
    return;
  }
  long long r = gid - nL;
  if (r < nLb) {
    const int job = (int)(r >> 8), n = (int)(r & 255);
    const int l = job + 1;
    if (l == 5) return;                         // from the per-frame sums below
    float s = 0.f;
    for (int sl = 0; sl < pl.SL; ++sl) s += partials[pl.off_L(job, sl) + 65536 + n];
    G.g[S2L_P_PTS0_W + 2 * l + 1][n] = s;
    return;
  }
  r -= nLb;
  if (r < nM) {
    const int which = (int)(r >> 14), e = (int)(r & 16383);
    float s = 0.f;
    for (int f = 0; f < pl.F; ++f)
      for (int sl = 0; sl < pl.SM; ++sl) s += partials[pl.off_M(which, f, sl) + e];
    red[(which ? R_M5 : R_M0) + e] = s;
    return;
  }
  r -= nM;
  if (r < nS) {
    const int which = (int)(r / (pl.F * 256)), f = (int)((r / 256) % pl.F), n = (int)(r & 255);
    float s = 0.f;
    for (int sl = 0; sl < pl.SM; ++sl) s += partials[pl.off_M(which, f, sl) + 16384 + n];
    red[R_S0 + ((long long)which * pl.F + f) * 256 + n] = s;
    return;
  }
  r -= nS;
  if (r < nO) {
    const int c = (int)(r >> 8), k = (int)(r & 255);
    if (c >= out_ch) return;
    float s = 0.f;
    for (int sl = 0; sl < pl.SO; ++sl) s += partials[pl.off_O(sl) + k * 16 + c];
    G.g[S2L_P_OUT_W][c * 256 + k] = s;
    return;
  }
  r -= nO;
  if (r < 4) {
    // d bias of output_linear = sum of dOut over all points (per-warp partial sums left by the dgrad kernel's prologue)
    const int c = (int)r;
    if (c >= out_ch) return;
    float s = 0.f;
    for (int i = 0; i < dbout_rows; ++i) s += dbout_part[i * 4 + c];
    G.g[S2L_P_OUT_B][c] = s;
  }
}

// ---- kernel 2: chain rule through fold0 = W0 Wuv and fold5 = W5a Wuvs  (one thread per output element)
__global__ void __launch_bounds__(256) wg_chain_kernel(const uint8_t* __restrict__ blob, Layout L, int F, int E, float* __restrict__ red,
                                                       const float* __restrict__ frame_bias, GradPtrs G) {
  const float* Fp = reinterpret_cast<const float*>(blob + L.off_fp32);
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nW = 2ll * 65536, nU = 2ll * 256 * E, nG = 2ll * F * 256, nB = 2ll * 256;
  if (gid < nW) {
    // dW0[i][j] = sum_e M0[i][e] Wuv[j][e] + sum_f S0_f[i] c_f[j];   same for W5[:, :256] with M5, Wuvs, S5, cs
    const int which = (int)(gid >> 16), i = (int)((gid >> 8) & 255), j = (int)(gid & 255);
    const float* M = red + (which ? R_M5 : R_M0) + i * 64;
    const float* WuT = Fp + (which ? F_UVS_WT : F_UV_WT);                   // [e][j]
    float s = 0.f;
    for (int e = 0; e < E; ++e) s = fmaf(M[e], WuT[e * 256 + j], s);
    for (int f = 0; f < F; ++f) s = fmaf(red[R_S0 + ((long long)which * F + f) * 256 + i], frame_bias[(size_t)f * 1024 + which * 256 + j], s);
    if (which) G.g[S2L_P_PTS0_W + 10][i * 512 + j] = s;
    else G.g[S2L_P_PTS0_W][i * 256 + j] = s;
    return;
  }
  long long r = gid - nW;
  if (r < nU) {
    // dWuv[j][e] = sum_i W0[i][j] M0[i][e]
    const int which = (int)(r / (256 * E)), j = (int)((r / E) % 256), e = (int)(r % E);
    const float* WT = Fp + f_pts_off(which ? 5 : 0) + j * 256;              // W^T row j: W[i][j] over i
    const float* M = red + (which ? R_M5 : R_M0) + e;
    float s = 0.f;
    for (int i = 0; i < 256; ++i) s = fmaf(WT[i], M[i * 64], s);
    G.g[which ? S2L_P_FC_UV_SKIP_W : S2L_P_FC_UV_W][j * E + e] = s;
    return;
  }
  r -= nU;
  if (r < nG) {
    // g_f[j] = sum_i W0[i][j] S0_f[i]   (gradient of the per-frame constant c_f)
    const int which = (int)(r / (F * 256)), f = (int)((r / 256) % F), j = (int)(r & 255);
    const float* WT = Fp + f_pts_off(which ? 5 : 0) + j * 256;
    const float* Sv = red + R_S0 + ((long long)which * F + f) * 256;
    float s = 0.f;
    for (int i = 0; i < 256; ++i) s = fmaf(WT[i], Sv[i], s);
    red[R_S0 + ((long long)(2 + which) * F + f) * 256 + j] = s;
    return;
  }
  r -= nG;
  if (r < nB) {
    // b0 / b5 gradients = column sums of dPre0 / dPre5 over all frames
    const int which = (int)(r >> 8), n = (int)(r & 255);
    float s = 0.f;
    for (int f = 0; f < F; ++f) s += red[R_S0 + ((long long)which * F + f) * 256 + n];
    G.g[S2L_P_PTS0_W + (which ? 11 : 1)][n] = s;
  }
}

// ---- kernel 3: everything that hangs off g0_f / g5_f: fc_audio, fc_time (+ skip twins), their biases, fc_uv biases, d latent
__global__ void __launch_bounds__(256) wg_frame_terms_kernel(const uint8_t* __restrict__ blob, Layout L, int F, const float* __restrict__ red,
                                                             const float* __restrict__ latent, const long long* __restrict__ frame_idx,
                                                             GradPtrs G) {
  const float* C = reinterpret_cast<const float*>(blob + L.off_const);
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nA = 2ll * 256 * 64, nT = 2ll * 256 * 20, nB = 2ll * 256, nL = (long long)F * 64;
  auto gvec = [&](int which, int f) { return red + R_S0 + ((long long)(2 + which) * F + f) * 256; };
  if (gid < nA) {
    const int which = (int)(gid / (256 * 64)), j = (int)((gid / 64) % 256), k = (int)(gid & 63);
    float s = 0.f;
    for (int f = 0; f < F; ++f) s = fmaf(gvec(which, f)[j], latent[f * 64 + k], s);
    G.g[which ? S2L_P_FC_AUDIO_SKIP_W : S2L_P_FC_AUDIO_W][j * 64 + k] = s;
    return;
  }
  long long r = gid - nA;
  if (r < nT) {
    const int which = (int)(r / (256 * 20)), j = (int)((r / 20) % 256), k = (int)(r % 20);
    float s = 0.f;
    for (int f = 0; f < F && frame_idx; ++f) {                                      // no time input: the term is absent, gradient 0
      const float ang = __fmul_rn((float)frame_idx[f], C[C_DIV + (k >> 1)]);      // tf_nerf.py:439-440
      s = fmaf(gvec(which, f)[j], (k & 1) ? cosf(ang) : sinf(ang), s);
    }
    G.g[which ? S2L_P_FC_TIME_SKIP_W : S2L_P_FC_TIME_W][j * 20 + k] = s;
    return;
  }
  r -= nT;
  if (r < nB) {
    const int which = (int)(r >> 8), j = (int)(r & 255);
    float s = 0.f;
    for (int f = 0; f < F; ++f) s += gvec(which, f)[j];
    const float st = frame_idx ? s : 0.f;
    if (which) {
      G.g[S2L_P_FC_UV_SKIP_B][j] = s; G.g[S2L_P_FC_AUDIO_SKIP_B][j] = s; G.g[S2L_P_FC_TIME_SKIP_B][j] = st;
    } else {
      G.g[S2L_P_FC_UV_B][j] = s; G.g[S2L_P_FC_AUDIO_B][j] = s; G.g[S2L_P_FC_TIME_B][j] = st;
    }
    return;
  }
  r -= nB;
  if (r < nL * 32) {
    // d latent_f[k] = sum_j g0_f[j] Wa[j][k] + g5_f[j] Was[j][k]: one warp per (f, k), lanes stride j (coalesced rows of the
    // transposed weights; a thread per output walked them 1 KB apart), fixed-order shuffle reduction
    const int lane = (int)(r & 31), o = (int)(r >> 5);
    const int f = o >> 6, k = o & 63;
    const float* g0 = gvec(0, f);
    const float* g5 = gvec(1, f);
    float s = 0.f;
#pragma unroll
    for (int j = lane; j < 256; j += 32) s = fmaf(g0[j], C[C_FCA_WT + k * 256 + j], fmaf(g5[j], C[C_FCAS_WT + k * 256 + j], s));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) G.d_latent[f * 64 + k] = s;
  }
}

struct TrainLayout {
  long long rows_total;
  size_t off_h, off_pe, off_dpre, off_dout, off_part, off_red, off_dbout, off_mask, total;
};
static TrainLayout train_layout_pts(long long P, int n_frames, int sms) {
  TrainLayout t{};
  const long long tiles = (P + 127) / 128;
  t.rows_total = tiles * 128 * (n_frames > 0 ? n_frames : 0);
  const WgPlan pl = make_wg_plan(n_frames, tiles, sms);
  auto al = [](size_t x) { return (x + 1023) & ~size_t(1023); };
  size_t o = 0;
  t.off_h = o;    o = al(o + (size_t)8 * t.rows_total * 256 * 2);
  t.off_pe = o;   o = al(o + (size_t)t.rows_total * 64 * 2);
  t.off_dpre = o; o = al(o + (size_t)8 * t.rows_total * 256 * 2);
  t.off_dout = o; o = al(o + (size_t)t.rows_total * 16 * 2);
  t.off_part = o; o = al(o + (size_t)pl.total_floats() * 4);
  t.off_red = o;  o = al(o + (size_t)red_floats(n_frames) * 4);
  t.off_dbout = o; o = al(o + (size_t)kDboutRows * 4 * 4);
  t.off_mask = o; o = al(o + (size_t)8 * t.rows_total * 8 * 4);      // 8 layers x 8 words per row
  t.total = o;
  return t;
}
static TrainLayout train_layout(const S2LGeom& g, int sms) {
  return train_layout_pts((long long)g.height * g.width * 4, g.n_frames, sms);
}
static int sm_count() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 148;
}
static int check_train_geom(const S2LGeom* g, const char* who) {
  if (!g) { set_error("%s: geom is null", who); return 1; }
  if (g->pts_mode != S2L_PTS_GRID_ENS4 || g->uv_dims != 2 || g->out_ch != 3 || g->n_frames < 0 || g->height < 0 || g->width < 0) {
    set_error("%s: the training render is the live 4-tap mode (pts_mode GRID_ENS4, uv_dims 2, out_ch 3)", who);
    return 2;
  }
  return 0;
}
static PointSrc train_src(const S2LGeom& g) {
  PointSrc s{};
  s.mode = S2L_PTS_GRID_ENS4;
  s.H = g.height;
  s.W = g.width;
  s.S = 1;
  s.Sc = 1;
  s.uv_dims = 2;
  s.eps = g.eps_shift;
  s.eps_pf = g.eps_per_frame;
  s.P = (long long)g.height * g.width * 4;
  s.R = g.height * g.width;
  s.step_w = g.width > 1 ? 1.0f / (float)(g.width - 1) : 0.f;
  s.step_h = g.height > 1 ? 1.0f / (float)(g.height - 1) : 0.f;
  return s;
}

}  // namespace s2l

using namespace s2l;

extern "C" size_t s2l_train_workspace_bytes(const S2LGeom* geom) {
  if (!geom || geom->n_frames < 0 || geom->height < 0 || geom->width < 0) return 0;
  return train_layout(*geom, sm_count()).total;
}

extern "C" int32_t s2l_train_fwd(const void* blob, const S2LGeom* geom, const float* latent, const int64_t* frame_idx, float* rgb,
                                 float* frame_bias, void* workspace, void* stream) {
  if (int e = check_train_geom(geom, "s2l_train_fwd")) return e;
  if (geom->n_frames == 0 || geom->height * geom->width == 0) return 0;
  if (!blob || !latent || !frame_idx || !rgb || !frame_bias || !workspace) { set_error("s2l_train_fwd: null argument"); return 1; }
  const TrainLayout t = train_layout(*geom, sm_count());
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  int rc = s2l_latent_bias_fwd(blob, latent, 64, frame_idx, frame_bias, geom->n_frames, stream);
  if (rc) return rc;
  return launch_mlp_tc_train(blob, train_src(*geom), geom->n_frames, frame_bias, rgb, reinterpret_cast<__nv_bfloat16*>(ws + t.off_h),
                             reinterpret_cast<__nv_bfloat16*>(ws + t.off_pe), reinterpret_cast<uint32_t*>(ws + t.off_mask),
                             reinterpret_cast<cudaStream_t>(stream));
}

// dgrad -> wgrad -> slab reduction -> fold chain rule -> per-frame terms, shared by the render and the rows entry points
static int train_backward(const void* blob, const PointSrc& src, int F, const TrainLayout& t, const float* d_rgb, const float* d_out_rows,
                          const float* latent, const int64_t* frame_idx, const float* frame_bias, void* workspace,
                          float* const* grads_host, float* d_latent, cudaStream_t st, const char* who) {
  GradPtrs G{};
  for (int i = S2L_P_FC_UV_W; i < S2L_NUM_PARAMS; ++i) {
    if (!grads_host[i]) { set_error("%s: gradient buffer %d is null", who, i); return 3; }
    G.g[i] = grads_host[i];
  }
  G.d_latent = d_latent;
  const int sms = sm_count();
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  TrainBufs B{};
  B.h = reinterpret_cast<__nv_bfloat16*>(ws + t.off_h);
  B.pe = reinterpret_cast<__nv_bfloat16*>(ws + t.off_pe);
  B.mask = reinterpret_cast<const uint32_t*>(ws + t.off_mask);
  B.dpre = reinterpret_cast<__nv_bfloat16*>(ws + t.off_dpre);
  B.dout16 = reinterpret_cast<__nv_bfloat16*>(ws + t.off_dout);
  B.partials = reinterpret_cast<float*>(ws + t.off_part);
  B.dbout_part = reinterpret_cast<float*>(ws + t.off_dbout);
  B.rows_total = t.rows_total;
  if (cudaMemsetAsync(B.dbout_part, 0, (size_t)kDboutRows * 4 * 4, st) != cudaSuccess) { set_error("%s: cudaMemsetAsync failed", who); return 5; }
  float* red = reinterpret_cast<float*>(ws + t.off_red);
  const long long tiles = (src.P + 127) / 128;
  const WgPlan pl = make_wg_plan(F, tiles, sms);
  int rc;
  if ((rc = launch_dgrad_tc(blob, src, F, d_rgb, B, st, d_out_rows))) return rc;
  if ((rc = launch_wgrad_tc(pl, B, st))) return rc;
  const Layout L = blob_layout();
  const int E = pe_dim(2);
  {
    const long long n = 7ll * 65536 + 7ll * 256 + 2ll * 16384 + 2ll * F * 256 + 256ll * 4 + 4;
    wg_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pl, B.partials, G, red, B.dbout_part, kDboutRows, 3);
    if (!check_launch("wg_reduce_kernel")) return 5;
  }
  {
    const long long n = 2ll * 65536 + 2ll * 256 * E + 2ll * F * 256 + 512;
    wg_chain_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint8_t*>(blob), L, F, E, red, frame_bias, G);
    if (!check_launch("wg_chain_kernel")) return 5;
  }
  {
    const long long n = 2ll * 256 * 64 + 2ll * 256 * 20 + 512 + (long long)F * 64 * 32;
    wg_frame_terms_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint8_t*>(blob), L, F, red, latent,
                                                                       reinterpret_cast<const long long*>(frame_idx), G);
    if (!check_launch("wg_frame_terms_kernel")) return 5;
  }
  return 0;
}

extern "C" int32_t s2l_train_bwd(const void* blob, const S2LGeom* geom, const float* d_rgb, const float* latent, const int64_t* frame_idx,
                                 const float* frame_bias, void* workspace, float* const* grads_host, float* d_latent, void* stream) {
  if (int e = check_train_geom(geom, "s2l_train_bwd")) return e;
  if (geom->n_frames == 0 || geom->height * geom->width == 0) return 0;
  if (!blob || !d_rgb || !latent || !frame_idx || !frame_bias || !workspace || !grads_host || !d_latent) { set_error("s2l_train_bwd: null argument"); return 1; }
  return train_backward(blob, train_src(*geom), geom->n_frames, train_layout(*geom, sm_count()), d_rgb, nullptr, latent, frame_idx, frame_bias,
                        workspace, grads_host, d_latent, reinterpret_cast<cudaStream_t>(stream), "s2l_train_bwd");
}

// ---- the per-call rows contract on the tensor-core training kernels (bf16): rows that share ONE latent
static PointSrc rows_src(const float* x, long long n_rows) {
  PointSrc s{};
  s.mode = S2L_PTS_EXPLICIT;
  s.uv_dims = 2;
  s.S = s.Sc = 1;
  s.P = n_rows;
  s.pts = x;
  s.pts_stride = 2 + kLatent;
  return s;
}
extern "C" size_t s2l_train_rows_workspace_bytes(int64_t n_rows) {
  return n_rows > 0 ? train_layout_pts(n_rows, 1, sm_count()).total : 0;
}
extern "C" int32_t s2l_train_rows_fwd(const void* blob, const float* x, int64_t n_rows, const int64_t* time_idx_dev, float* out,
                                      float* frame_bias, void* workspace, void* stream) {
  if (n_rows < 0) { set_error("s2l_train_rows_fwd: negative n_rows"); return 2; }
  if (n_rows == 0) return 0;
  if (!blob || !x || !out || !frame_bias || !workspace) { set_error("s2l_train_rows_fwd: null argument"); return 1; }
  const TrainLayout t = train_layout_pts(n_rows, 1, sm_count());
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  int rc = s2l_latent_bias_fwd(blob, x + 2, 2 + kLatent, time_idx_dev, frame_bias, 1, stream);
  if (rc) return rc;
  return launch_mlp_tc_train(blob, rows_src(x, n_rows), 1, frame_bias, nullptr, reinterpret_cast<__nv_bfloat16*>(ws + t.off_h),
                             reinterpret_cast<__nv_bfloat16*>(ws + t.off_pe), reinterpret_cast<uint32_t*>(ws + t.off_mask),
                             reinterpret_cast<cudaStream_t>(stream), out);
}
extern "C" int32_t s2l_train_rows_bwd(const void* blob, const float* d_out, const float* x, int64_t n_rows, const int64_t* time_idx_dev,
                                      const float* frame_bias, void* workspace, float* const* grads_host, float* d_latent, void* stream) {
  if (n_rows < 0) { set_error("s2l_train_rows_bwd: negative n_rows"); return 2; }
  if (n_rows == 0) return 0;
  if (!blob || !d_out || !x || !frame_bias || !workspace || !grads_host || !d_latent) { set_error("s2l_train_rows_bwd: null argument"); return 1; }
  return train_backward(blob, rows_src(x, n_rows), 1, train_layout_pts(n_rows, 1, sm_count()), nullptr, d_out, x + 2, time_idx_dev, frame_bias,
                        workspace, grads_host, d_latent, reinterpret_cast<cudaStream_t>(stream), "s2l_train_rows_bwd");
}
