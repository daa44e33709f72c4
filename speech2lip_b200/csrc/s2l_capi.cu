// C-ABI glue: argument validation, error reporting, launch bookkeeping and the whole-path entry point.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <utility>
#include <vector>
#include "s2l_common.cuh"
#include "s2l_points.cuh"

namespace s2l {

static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
bool check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return false;
  }
  g_launches += 1;
  return true;
}

// ---- bench.py's live kernel timing: CUDA events around every fused-MLP launch, on the launching stream
static bool g_prof_on = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_pool;
static size_t g_prof_used = 0;
void profile_mark(cudaStream_t st, int which) {
  if (!g_prof_on) return;
  if (which == 0) {
    if (g_prof_used == g_prof_pool.size()) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
      g_prof_pool.emplace_back(a, b);
    }
    cudaEventRecord(g_prof_pool[g_prof_used].first, st);
  } else if (g_prof_used < g_prof_pool.size()) {
    cudaEventRecord(g_prof_pool[g_prof_used].second, st);
    ++g_prof_used;
  }
}

int launch_mlp_fp32(const void* blob, const PointSrc& src, int n_frames, const float* frame_bias, float* out,
                    int out_ch, cudaStream_t st, const float4* fix_carry = nullptr, float* fix_rgb = nullptr,
                    float* patch_raw = nullptr);
int launch_tile_scan(const int* count, int F, int Sc, int* ts128, int* ts64, cudaStream_t st);
int launch_flag_last(const float* raw, int F, int R, int S, float thr, const float* auto_thr, int* count, int* rays, cudaStream_t st);
int launch_mlp_fp32_rows(const void* blob, const float* x, long long n_rows, long long time_idx, int has_time,
                         float* out, float* save, int uv_dims, int out_ch, cudaStream_t st, const Gate* gate = nullptr,
                         const long long* time_idx_dev = nullptr);
int launch_mlp_bwd_rows(const void* blob, const float* d_out, const float* acts, long long n_rows, float* dsave,
                        int out_ch, cudaStream_t st);
int launch_embed(const float* x, long long n_rows, int row_stride, int uv_dims, float* pe, cudaStream_t st);
int launch_mlp_tc(const void* blob, const PointSrc& src, int n_frames, const float* frame_bias, float* out, int out_ch,
                  int npass, cudaStream_t st, const TcEpi* epi = nullptr, const Gate* gate = nullptr);

static long long points_per_frame(const S2LGeom& g) {
  switch (g.pts_mode) {
    case S2L_PTS_GRID: return (long long)g.height * g.width;
    case S2L_PTS_GRID_ENS4: return (long long)g.height * g.width * 4;
    case S2L_PTS_RAYS: return (long long)g.height * g.width * g.n_samples;
    case S2L_PTS_EXPLICIT: return g.pts_per_frame;
    default: return -1;
  }
}

static int validate_geom(const S2LGeom* g, const char* who) {
  if (!g) { set_error("%s: geom is null", who); return 1; }
  if (g->n_frames < 0 || g->height < 0 || g->width < 0) { set_error("%s: negative geometry", who); return 2; }
  if (g->uv_dims != 2 && g->uv_dims != 3) { set_error("%s: uv_dims must be 2 or 3 (got %d)", who, g->uv_dims); return 2; }
  if (g->out_ch < 1 || g->out_ch > 4) { set_error("%s: out_ch must be in 1..4 (got %d)", who, g->out_ch); return 2; }
  if (g->pts_mode < 0 || g->pts_mode > S2L_PTS_EXPLICIT) { set_error("%s: unknown pts_mode %d", who, g->pts_mode); return 2; }
  if ((g->pts_mode == S2L_PTS_GRID || g->pts_mode == S2L_PTS_GRID_ENS4) && g->uv_dims != 2) {
    set_error("%s: grid modes need the uv_dims=2 model (got %d)", who, g->uv_dims);
    return 2;
  }
  if (g->pts_mode == S2L_PTS_RAYS && (g->uv_dims != 3 || g->n_samples < 1)) {
    set_error("%s: ray mode needs the uv_dims=3 model and n_samples >= 1 (uv_dims=%d, n_samples=%d)", who, g->uv_dims, g->n_samples);
    return 2;
  }
  if (g->pts_mode == S2L_PTS_EXPLICIT && g->pts_per_frame < 0) { set_error("%s: negative pts_per_frame", who); return 2; }
  return 0;
}

static PointSrc make_src(const S2LGeom& g, const float* pts, const float* ro, const float* rd, const float* z) {
  PointSrc s{};
  s.mode = g.pts_mode;
  s.H = g.height;
  s.W = g.width;
  s.S = g.n_samples > 0 ? g.n_samples : 1;
  s.uv_dims = g.uv_dims;
  s.z_per_ray = g.z_per_ray;
  s.rays_shared = g.rays_per_frame_shared;
  s.eps = g.eps_shift;
  s.eps_pf = (g.pts_mode == S2L_PTS_GRID_ENS4) ? g.eps_per_frame : nullptr;
  s.P = points_per_frame(g);
  s.R = g.height * g.width;
  s.step_w = g.width > 1 ? 1.0f / (float)(g.width - 1) : 0.f;
  s.step_h = g.height > 1 ? 1.0f / (float)(g.height - 1) : 0.f;
  s.s0 = 0;
  s.Sc = s.S;
  s.pts = pts;
  s.rays_o = ro;
  s.rays_d = rd;
  s.z = z;
  return s;
}

}  // namespace s2l

using namespace s2l;

extern "C" const char* s2l_last_error(void) { return g_err; }
extern "C" int32_t s2l_abi_version(void) { return S2L_ABI_VERSION; }
extern "C" int64_t s2l_launch_count(int32_t reset) {
  const long long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

extern "C" int32_t s2l_sizeof_geom(void) { return (int32_t)sizeof(S2LGeom); }
extern "C" void s2l_profile_enable(int32_t on) {
  g_prof_on = on != 0;
  if (on) g_prof_used = 0;
}
extern "C" double s2l_profile_mlp_ms(int32_t* n_launches) {
  double ms = 0.0;
  for (size_t i = 0; i < g_prof_used; ++i) {
    float t = 0.f;
    cudaEventSynchronize(g_prof_pool[i].second);
    if (cudaEventElapsedTime(&t, g_prof_pool[i].first, g_prof_pool[i].second) == cudaSuccess) ms += t;
  }
  if (n_launches) *n_launches = (int32_t)g_prof_used;
  g_prof_used = 0;
  return ms;
}

extern "C" int32_t s2l_mlp_fwd(const void* blob, const S2LGeom* geom, const float* frame_bias, const float* pts,
                               const float* rays_o, const float* rays_d, const float* z_vals, float* raw_out,
                               int32_t precision, void* stream) {
  if (int e = validate_geom(geom, "s2l_mlp_fwd")) return e;
  if (!blob || !frame_bias || !raw_out) { set_error("s2l_mlp_fwd: null blob/frame_bias/raw_out"); return 1; }
  if (geom->pts_mode == S2L_PTS_EXPLICIT && !pts && geom->pts_per_frame > 0) { set_error("s2l_mlp_fwd: EXPLICIT mode needs pts"); return 1; }
  if (geom->pts_mode == S2L_PTS_RAYS && (!rays_o || !rays_d || !z_vals)) { set_error("s2l_mlp_fwd: RAYS mode needs rays_o, rays_d, z_vals"); return 1; }
  const PointSrc src = make_src(*geom, pts, rays_o, rays_d, z_vals);
  if (src.P == 0 || geom->n_frames == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (precision) {
    case S2L_PREC_FP32: return launch_mlp_fp32(blob, src, geom->n_frames, frame_bias, raw_out, geom->out_ch, st);
    case S2L_PREC_BF16X3: return launch_mlp_tc(blob, src, geom->n_frames, frame_bias, raw_out, geom->out_ch, 3, st);
    case S2L_PREC_BF16X1: return launch_mlp_tc(blob, src, geom->n_frames, frame_bias, raw_out, geom->out_ch, 1, st);
    case S2L_PREC_FP16F8: return launch_mlp_tc(blob, src, geom->n_frames, frame_bias, raw_out, geom->out_ch, 2, st);
    default: set_error("s2l_mlp_fwd: unknown precision %d", precision); return 2;
  }
}

extern "C" int32_t s2l_rgb_forward_rows(const void* blob, const float* x, int64_t n_rows, int64_t time_idx,
                                        int32_t has_time, float* out, int32_t uv_dims, int32_t out_ch, void* stream) {
  if (!blob || (!x && n_rows > 0) || (!out && n_rows > 0)) { set_error("s2l_rgb_forward_rows: null argument"); return 1; }
  if (n_rows < 0) { set_error("s2l_rgb_forward_rows: negative n_rows"); return 2; }
  if ((uv_dims != 2 && uv_dims != 3) || out_ch < 1 || out_ch > 4) { set_error("s2l_rgb_forward_rows: unsupported dims uv_dims=%d out_ch=%d", uv_dims, out_ch); return 2; }
  return launch_mlp_fp32_rows(blob, x, n_rows, time_idx, has_time, out, nullptr, uv_dims, out_ch, reinterpret_cast<cudaStream_t>(stream));
}

// Drop-in rgb_forward without host synchronisation (include/speech2lip_b200.h): compare kernel -> device flag; the
// constant-latent tensor-core path and the general per-row fp32 path are BOTH enqueued, each gated on the flag.
extern "C" size_t s2l_rgb_forward_auto_scratch_bytes(void) { return 4 * 256 * sizeof(float) + 256; }
extern "C" int32_t s2l_rgb_forward_auto(const void* blob, const float* x, int64_t n_rows, const int64_t* time_idx_dev, float* out,
                                        int32_t uv_dims, int32_t out_ch, int32_t precision, void* scratch, void* stream) {
  if (!blob || (n_rows > 0 && (!x || !out)) || !scratch) { set_error("s2l_rgb_forward_auto: null argument"); return 1; }
  if (n_rows < 0) { set_error("s2l_rgb_forward_auto: negative n_rows"); return 2; }
  if ((uv_dims != 2 && uv_dims != 3) || out_ch < 1 || out_ch > 4) { set_error("s2l_rgb_forward_auto: unsupported dims uv_dims=%d out_ch=%d", uv_dims, out_ch); return 2; }
  const int np = precision == S2L_PREC_BF16X3 ? 3 : precision == S2L_PREC_FP16F8 ? 2 : precision == S2L_PREC_BF16X1 ? 1 : 0;
  if (n_rows == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (np == 0)      // exact precision requested: the general kernel serves both cases
    return launch_mlp_fp32_rows(blob, x, n_rows, 0, time_idx_dev ? 1 : 0, out, nullptr, uv_dims, out_ch, st, nullptr,
                                reinterpret_cast<const long long*>(time_idx_dev));
  float* bias = reinterpret_cast<float*>(scratch);
  int32_t* flag = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(scratch) + 4 * 256 * sizeof(float));
  const int stride = uv_dims + kLatent;
  int rc = s2l_rows_differ(x, n_rows, stride, uv_dims, kLatent, flag, stream);
  if (rc) return rc;
  if ((rc = s2l_latent_bias_fwd(blob, x + uv_dims, stride, time_idx_dev, bias, 1, stream))) return rc;
  PointSrc src{};
  src.mode = S2L_PTS_EXPLICIT;
  src.uv_dims = uv_dims;
  src.S = src.Sc = 1;
  src.P = n_rows;
  src.pts = x;
  src.pts_stride = stride;
  const Gate g_const{flag, 0}, g_rows{flag, 1};
  if ((rc = launch_mlp_tc(blob, src, 1, bias, out, out_ch, np, st, nullptr, &g_const))) return rc;
  return launch_mlp_fp32_rows(blob, x, n_rows, 0, time_idx_dev ? 1 : 0, out, nullptr, uv_dims, out_ch, st, &g_rows,
                              reinterpret_cast<const long long*>(time_idx_dev));
}

extern "C" int32_t s2l_rgb_forward_rows_train(const void* blob, const float* x, int64_t n_rows, int64_t time_idx,
                                              int32_t has_time, float* out, float* acts, int32_t uv_dims, int32_t out_ch,
                                              void* stream) {
  if (!blob || (n_rows > 0 && (!x || !out || !acts))) { set_error("s2l_rgb_forward_rows_train: null argument"); return 1; }
  if (n_rows < 0) { set_error("s2l_rgb_forward_rows_train: negative n_rows"); return 2; }
  if ((uv_dims != 2 && uv_dims != 3) || out_ch < 1 || out_ch > 4) { set_error("s2l_rgb_forward_rows_train: unsupported dims uv_dims=%d out_ch=%d", uv_dims, out_ch); return 2; }
  return launch_mlp_fp32_rows(blob, x, n_rows, time_idx, has_time, out, acts, uv_dims, out_ch, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int32_t s2l_mlp_bwd_rows(const void* blob, const float* d_out, const float* acts, int64_t n_rows, float* dsave,
                                    int32_t out_ch, void* stream) {
  if (!blob || (n_rows > 0 && (!d_out || !acts || !dsave))) { set_error("s2l_mlp_bwd_rows: null argument"); return 1; }
  if (n_rows < 0 || out_ch < 1 || out_ch > 4) { set_error("s2l_mlp_bwd_rows: bad n_rows/out_ch"); return 2; }
  return launch_mlp_bwd_rows(blob, d_out, acts, n_rows, dsave, out_ch, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int32_t s2l_embed_fwd(const float* x, int64_t n_rows, int32_t row_stride, int32_t uv_dims, float* pe, void* stream) {
  if (n_rows > 0 && (!x || !pe)) { set_error("s2l_embed_fwd: null argument"); return 1; }
  if (n_rows < 0 || (uv_dims != 2 && uv_dims != 3) || row_stride < uv_dims) { set_error("s2l_embed_fwd: bad arguments"); return 2; }
  return launch_embed(x, n_rows, row_stride, uv_dims, pe, reinterpret_cast<cudaStream_t>(stream));
}

// ---- whole-path orchestration ------------------------------------------------------------------------------------
static int npass_of(int precision) {
  return precision == S2L_PREC_BF16X3 ? 3 : precision == S2L_PREC_BF16X1 ? 1 : precision == S2L_PREC_FP16F8 ? 2 : 0;
}

// What a s2l_render_frames call does and where its scratch sections live.
struct RenderPlan {
  bool tc, fused, fix;
  int chunks, Sc;
  float fix_thr, term_thr;
  size_t off_bias, off_carry, off_list[2], off_counts, off_ts128, off_ts64, off_raw, total;
};
static RenderPlan make_plan(const S2LGeom& g, int precision, bool want_aux) {
  RenderPlan p{};
  const long long P = points_per_frame(g);
  const size_t F = (size_t)(g.n_frames > 0 ? g.n_frames : 0);
  const size_t R = (size_t)g.height * (size_t)g.width;
  p.tc = npass_of(precision) != 0;
  p.chunks = 1;
  p.Sc = g.n_samples;
  const bool rays = g.pts_mode == S2L_PTS_RAYS;
  auto chunk_ok = [](int sc) { return sc >= 4 && sc <= 128 && (sc & 3) == 0 && 128 % sc == 0; };
  if (rays && p.tc && !want_aux && g.sample_chunks > 1 && g.n_samples % g.sample_chunks == 0 && chunk_ok(g.n_samples / g.sample_chunks)) {
    p.chunks = g.sample_chunks;
    p.Sc = g.n_samples / g.sample_chunks;
  }
  p.fused = p.tc && ((rays && !want_aux && g.out_ch == 4 && chunk_ok(p.Sc)) || (g.pts_mode == S2L_PTS_GRID_ENS4 && g.out_ch == 3));
  p.fix_thr = g.fix_thr == 0.f ? -1.f : g.fix_thr;      // -1 = the automatic threshold in the blob (META[3]); geom.fix_thr < 0 = off
  p.fix = rays && p.tc && precision != S2L_PREC_BF16X1 && g.fix_thr >= 0.f;      // bf16x1 is a preview mode: no parity claim to protect
  p.term_thr = g.term_thr;
  auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
  size_t o = 0;
  p.off_bias = o;   o = al(o + F * 4 * 256 * sizeof(float));
  const bool lists = rays && p.tc && (p.fix || p.chunks > 1);
  // the counters come first so that their offset does not depend on which sections follow (s2l_render_counts_offset)
  p.off_counts = o; if (rays) o = al(o + (size_t)(p.chunks + 1) * F * sizeof(int));
  p.off_ts128 = o;  if (lists) o = al(o + (F + 1) * sizeof(int));
  p.off_ts64 = o;   if (lists) o = al(o + (F + 1) * sizeof(int));
  p.off_carry = o;  if (lists && p.fused) o = al(o + F * R * sizeof(float4));
  p.off_list[0] = o; if (lists) o = al(o + F * R * sizeof(int));
  p.off_list[1] = o; if (lists && p.chunks > 1) o = al(o + F * R * sizeof(int));
  p.off_raw = o;
  const bool direct = (g.pts_mode == S2L_PTS_GRID && g.out_ch == 3);
  if (!p.fused && !direct && P > 0) o = al(o + F * (size_t)P * g.out_ch * sizeof(float));
  p.total = o + 256;
  return p;
}

extern "C" size_t s2l_render_scratch_bytes(const S2LGeom* g, int32_t precision, int32_t want_aux) {
  if (!g || points_per_frame(*g) < 0) return 0;
  return make_plan(*g, precision, want_aux != 0).total;
}
extern "C" size_t s2l_render_counts_offset(const S2LGeom* g) {
  if (!g) return 0;
  return make_plan(*g, S2L_PREC_FP16F8, false).off_counts;
}

extern "C" int32_t s2l_render_frames(const void* blob, const S2LGeom* geom, const float* audio, const int64_t* frame_idx,
                                     const float* rays_o, const float* rays_d, const float* z_vals, float* rgb,
                                     float* weights, float* depth, void* scratch, int32_t precision, void* stream) {
  if (int e = validate_geom(geom, "s2l_render_frames")) return e;
  if (geom->n_frames == 0 || points_per_frame(*geom) == 0) return 0;      /* empty batch: nothing to do */
  if (!blob || !audio || !rgb || !scratch) { set_error("s2l_render_frames: null blob/audio/rgb/scratch"); return 1; }
  if (geom->pts_mode == S2L_PTS_EXPLICIT) { set_error("s2l_render_frames: EXPLICIT points are served by s2l_mlp_fwd"); return 2; }
  if (geom->pts_mode == S2L_PTS_RAYS && geom->out_ch != 4) { set_error("s2l_render_frames: ray mode needs the out_ch=4 model"); return 2; }
  if (geom->pts_mode == S2L_PTS_GRID && geom->out_ch != 3) { set_error("s2l_render_frames: GRID mode needs the out_ch=3 model"); return 2; }
  if (geom->pts_mode == S2L_PTS_RAYS && (!rays_o || !rays_d || !z_vals)) { set_error("s2l_render_frames: RAYS mode needs rays_o, rays_d, z_vals"); return 1; }
  if (precision < S2L_PREC_FP32 || precision > S2L_PREC_FP16F8) { set_error("s2l_render_frames: unknown precision %d", precision); return 2; }
  const bool want_aux = weights || depth;
  const RenderPlan pl = make_plan(*geom, precision, want_aux);
  if (geom->pts_mode == S2L_PTS_RAYS && geom->sample_chunks > 1 && pl.chunks == 1) {
    set_error("s2l_render_frames: sample_chunks=%d needs a tensor-core precision, no weights/depth outputs and "
              "n_samples/sample_chunks in {4,8,...,128} (n_samples=%d)", geom->sample_chunks, geom->n_samples);
    return 2;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* sc = reinterpret_cast<uint8_t*>(scratch);
  float* bias = reinterpret_cast<float*>(sc + pl.off_bias);
  float* raw = reinterpret_cast<float*>(sc + pl.off_raw);
  float4* carry = reinterpret_cast<float4*>(sc + pl.off_carry);
  int* lists[2] = {reinterpret_cast<int*>(sc + pl.off_list[0]), reinterpret_cast<int*>(sc + pl.off_list[1])};
  int* counts = reinterpret_cast<int*>(sc + pl.off_counts);
  int* ts128 = reinterpret_cast<int*>(sc + pl.off_ts128);
  int* ts64 = reinterpret_cast<int*>(sc + pl.off_ts64);
  const int F = geom->n_frames;
  const int R = geom->height * geom->width;
  const int np = npass_of(precision);
  int rc = s2l_audio_encode_fwd(blob, audio, 0, frame_idx, nullptr, bias, F, geom->uv_dims, geom->out_ch, stream);
  if (rc) return rc;

  if (geom->pts_mode == S2L_PTS_GRID)
    return s2l_mlp_fwd(blob, geom, bias, nullptr, nullptr, nullptr, nullptr, rgb, precision, stream);

  if (geom->pts_mode == S2L_PTS_GRID_ENS4) {
    if (pl.fused) {
      const PointSrc src = make_src(*geom, nullptr, nullptr, nullptr, nullptr);
      TcEpi epi{};
      epi.mode = EPI_ENS4;
      epi.rgb = rgb;
      return launch_mlp_tc(blob, src, F, bias, nullptr, geom->out_ch, np, st, &epi);
    }
    rc = s2l_mlp_fwd(blob, geom, bias, nullptr, nullptr, nullptr, nullptr, raw, precision, stream);
    if (rc) return rc;
    return s2l_ensemble4_blend(raw, geom, rgb, stream);
  }

  // ---- RAYS
  if (cudaMemsetAsync(counts, 0, (size_t)(pl.chunks + 1) * F * sizeof(int), st) != cudaSuccess) {
    set_error("s2l_render_frames: cudaMemsetAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
    return 5;
  }
  // list of rays the fp32 kernel re-evaluates at their last sample
  PointSrc fixsrc = make_src(*geom, nullptr, rays_o, rays_d, z_vals);
  fixsrc.s0 = geom->n_samples - 1;
  fixsrc.Sc = 1;
  fixsrc.P = R;
  fixsrc.list_count = counts + (size_t)pl.chunks * F;
  fixsrc.tile_start = ts64;

  if (!pl.fused) {
    rc = s2l_mlp_fwd(blob, geom, bias, nullptr, rays_o, rays_d, z_vals, raw, precision, stream);
    if (rc) return rc;
    if (pl.fix) {
      int* fcount = counts + (size_t)pl.chunks * F;
      const float* auto_thr = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(blob) + blob_layout().off_meta) + 3;
      if ((rc = launch_flag_last(raw, F, R, geom->n_samples, pl.fix_thr, auto_thr, fcount, lists[0], st))) return rc;
      if ((rc = launch_tile_scan(fcount, F, 1, ts128, ts64, st))) return rc;
      fixsrc.list_rays = lists[0];
      if ((rc = launch_mlp_fp32(blob, fixsrc, F, bias, nullptr, 4, st, nullptr, nullptr, raw))) return rc;
    }
    return s2l_composite_fwd(raw, z_vals, geom->z_per_ray, rays_d, (long long)R * F, geom->rays_per_frame_shared ? R : 0,
                             geom->n_samples, rgb, weights, depth, stream);
  }

  // fused: C launches of Sc samples; launch c reads list c (c >= 1) and appends to list c + 1 (survivors) or, on the
  // last chunk, to the fp32 re-evaluation list
  for (int c = 0; c < pl.chunks; ++c) {
    PointSrc src = make_src(*geom, nullptr, rays_o, rays_d, z_vals);
    src.s0 = c * pl.Sc;
    src.Sc = pl.Sc;
    src.P = (long long)R * pl.Sc;
    if (c > 0) {
      src.list_count = counts + (size_t)c * F;
      src.list_rays = lists[(c - 1) & 1];
      src.tile_start = ts128;
      if ((rc = launch_tile_scan(src.list_count, F, pl.Sc, ts128, ts64, st))) return rc;
    }
    const bool last = (c == pl.chunks - 1);
    TcEpi epi{};
    epi.mode = EPI_COMPOSITE;
    epi.rgb = rgb;
    epi.carry = carry;
    epi.next_count = counts + (size_t)(c + 1) * F;
    epi.next_rays = lists[c & 1];
    epi.term_thr = pl.term_thr;
    epi.fix_thr = (last && pl.fix) ? pl.fix_thr : 0.f;
    if ((rc = launch_mlp_tc(blob, src, F, bias, nullptr, 4, np, st, &epi))) return rc;
  }
  if (pl.fix) {
    if ((rc = launch_tile_scan(fixsrc.list_count, F, 1, ts128, ts64, st))) return rc;
    fixsrc.list_rays = lists[(pl.chunks - 1) & 1];
    if ((rc = launch_mlp_fp32(blob, fixsrc, F, bias, nullptr, 4, st, carry, rgb, nullptr))) return rc;
  }
  return 0;
}
