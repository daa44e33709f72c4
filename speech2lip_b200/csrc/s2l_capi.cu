// C-ABI glue: argument validation, error reporting, launch bookkeeping and the whole-path entry point.
#include <cstdarg>
#include <cstdio>
#include <cstring>
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

int launch_mlp_fp32(const void* blob, const PointSrc& src, int n_frames, const float* frame_bias, float* out,
                    int out_ch, cudaStream_t st);
int launch_mlp_fp32_rows(const void* blob, const float* x, long long n_rows, long long time_idx, int has_time,
                         float* out, float* save, int uv_dims, int out_ch, cudaStream_t st);
int launch_mlp_bwd_rows(const void* blob, const float* d_out, const float* acts, long long n_rows, float* dsave,
                        int out_ch, cudaStream_t st);
int launch_embed(const float* x, long long n_rows, int row_stride, int uv_dims, float* pe, cudaStream_t st);
int launch_mlp_tc(const void* blob, const PointSrc& src, int n_frames, const float* frame_bias, float* out, int out_ch,
                  int npass, cudaStream_t st);

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

extern "C" size_t s2l_render_scratch_bytes(const S2LGeom* g) {
  if (!g) return 0;
  const long long P = points_per_frame(*g);
  if (P < 0) return 0;
  size_t bias = (size_t)g->n_frames * 4 * 256 * sizeof(float);
  size_t raw = (g->pts_mode == S2L_PTS_GRID && g->out_ch == 3) ? 0 : (size_t)g->n_frames * (size_t)P * g->out_ch * sizeof(float);
  return ((bias + 255) & ~size_t(255)) + raw + 256;
}

extern "C" int32_t s2l_render_frames(const void* blob, const S2LGeom* geom, const float* audio, const int64_t* frame_idx,
                                     const float* rays_o, const float* rays_d, const float* z_vals, float* rgb,
                                     float* weights, float* depth, void* scratch, int32_t precision, void* stream) {
  if (int e = validate_geom(geom, "s2l_render_frames")) return e;
  if (geom->n_frames == 0 || points_per_frame(*geom) == 0) return 0;      /* empty batch: nothing to do */
  if (!blob || !audio || !rgb || !scratch) { set_error("s2l_render_frames: null blob/audio/rgb/scratch"); return 1; }
  if (geom->pts_mode == S2L_PTS_EXPLICIT) { set_error("s2l_render_frames: EXPLICIT points are served by s2l_mlp_fwd"); return 2; }
  if (geom->pts_mode == S2L_PTS_RAYS && geom->out_ch != 4) { set_error("s2l_render_frames: ray mode needs the out_ch=4 model"); return 2; }
  float* bias = reinterpret_cast<float*>(scratch);
  const size_t bias_bytes = (((size_t)geom->n_frames * 4 * 256 * sizeof(float)) + 255) & ~size_t(255);
  float* raw = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(scratch) + bias_bytes);
  int rc = s2l_audio_encode_fwd(blob, audio, 0, frame_idx, nullptr, bias, geom->n_frames, geom->uv_dims, geom->out_ch, stream);
  if (rc) return rc;
  const bool direct = (geom->pts_mode == S2L_PTS_GRID && geom->out_ch == 3);
  rc = s2l_mlp_fwd(blob, geom, bias, nullptr, rays_o, rays_d, z_vals, direct ? rgb : raw, precision, stream);
  if (rc) return rc;
  if (geom->pts_mode == S2L_PTS_GRID) {
    if (!direct) { set_error("s2l_render_frames: GRID mode needs the out_ch=3 model"); return 2; }
    return 0;
  }
  if (geom->pts_mode == S2L_PTS_GRID_ENS4) return s2l_ensemble4_blend(raw, geom, rgb, stream);
  const long long R = (long long)geom->height * geom->width;
  return s2l_composite_fwd(raw, z_vals, geom->z_per_ray, rays_d, R * geom->n_frames,
                           geom->rays_per_frame_shared ? R : 0, geom->n_samples, rgb, weights, depth, stream);
}
