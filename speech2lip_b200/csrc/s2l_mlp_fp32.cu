// Exact-path fused MLP: CUDA-core fp32 FFMA, literal (unfolded) layer order of
// TalkingFace.rgb_forward (tf_nerf.py:225-285).  One persistent CTA per SM walks 64-point tiles;
// activations never leave shared memory, weights (W^T, [K][256] fp32) are streamed from L2 in 16-row
// chunks by 1-D bulk copies (UBLKCP) through a 3-stage mbarrier ring.
//
// Roles: this kernel is (a) the arithmetic reference on the GPU (true fp32), (b) the general
// rgb_forward contract with an arbitrary latent per row (ROWLAT), where the audio term cannot be
// hoisted into a per-frame bias.  The tensor-core kernel (s2l_mlp_tc.cu) is the throughput path.
#include "s2l_fp32_core.cuh"

namespace s2l {

template <bool ROWLAT>
__global__ void __launch_bounds__(256, 1) mlp_fp32_kernel(const __grid_constant__ Fp32Args a) {
  if (a.gate.flag && *a.gate.flag != a.gate.value) return;      // gated launch: the other implementation serves this call
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  float* X = reinterpret_cast<float*>(smem_raw);
  float* Y = X + 256 * TMP;
  float* PE = Y + 256 * TMP;                       // [64][TMP]
  float* LAT = PE + 64 * TMP;                      // [64][TMP] (ROWLAT only)
  float* wst = (ROWLAT ? LAT + 64 * TMP : LAT);    // NS * CHUNK_FLOATS, 16B aligned
  __shared__ uint64_t full[NS];
  __shared__ float bias0[256], biasS[256];
  __shared__ float tpe[kTimePE];

  const int tid = threadIdx.x;
  const int tn = tid & 31, m0 = (tid >> 5) * 8;
  const float* Fp = reinterpret_cast<const float*>(a.blob + a.L.off_fp32);
  const float* Cc = reinterpret_cast<const float*>(a.blob + a.L.off_const);
  const long long n_tiles = a.src.tile_start ? (long long)a.src.tile_start[a.n_frames] : a.tiles_per_frame * a.n_frames;
  const long long my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  Pipe ps{0, my_tiles * a.prog.n_chunks};
  const int E = pe_dim(a.src.uv_dims);
  const int e_chunks = (E + KC - 1) / KC;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0)
    for (long long c = 0; c < NS && c < ps.total; ++c) issue_chunk(a, wst, full, c);

  if (ROWLAT) {
    // per-call constant part of the first-layer / skip biases: b_uv + b_a + (Wt t + b_t)  (tf_nerf.py:252-258)
    if (tid < 10) {
      float s = 0.f, c = 1.f;
      if (a.has_time) {
        const float ang = __fmul_rn((float)(a.time_idx_dev ? a.time_idx_dev[0] : a.time_idx), Cc[C_DIV + tid]);
        s = sinf(ang);
        c = cosf(ang);
      }
      tpe[2 * tid] = s;
      tpe[2 * tid + 1] = c;
    }
    __syncthreads();
    float tt = 0.f, tts = 0.f;
    for (int k = 0; k < kTimePE; ++k) {
      tt = fmaf(Cc[C_FCT_WT + k * 256 + tid], tpe[k], tt);
      tts = fmaf(Cc[C_FCTS_WT + k * 256 + tid], tpe[k], tts);
    }
    float b0 = Cc[C_BIAS6 + 0 * 256 + tid] + Cc[C_BIAS6 + 1 * 256 + tid];
    float bs = Cc[C_BIAS6 + 3 * 256 + tid] + Cc[C_BIAS6 + 4 * 256 + tid];
    if (a.has_time) {
      b0 += tt + Cc[C_BIAS6 + 2 * 256 + tid];
      bs += tts + Cc[C_BIAS6 + 5 * 256 + tid];
    }
    bias0[tid] = b0;
    biasS[tid] = bs;
  }

  float acc[8][8];
  int fcur = 0;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    int f; long long p_base, Pf;
    tile_locate<TM>(a.src, a.tiles_per_frame, n_tiles, tile, fcur, f, p_base, Pf);
    // training forward: slot s of `save` is [F*P][256]; this tile's rows start at grow(s)
    const long long n_rows_total = (long long)a.n_frames * a.src.P;
    const int rows_valid = (int)((Pf - p_base) < TM ? (Pf - p_base) : TM);
    auto grow = [&](int slot) -> float* {
      return a.save ? a.save + ((size_t)slot * n_rows_total + (size_t)f * a.src.P + p_base) * 256 : nullptr;
    };
    // ---- stage the tile's inputs: positional encoding (tf_nerf.py:404-425) [+ latent rows]
    {
      const int m = tid & 63, part = tid >> 6;   // 4 threads per point, frequencies interleaved
      const long long p = p_base + m;
      float x[3] = {0.f, 0.f, 0.f};
      if (p < Pf) {
        if (ROWLAT) {
          const float* row = a.rows + p * (a.src.uv_dims + kLatent);
          for (int d = 0; d < a.src.uv_dims; ++d) x[d] = row[d];
        } else {
          gen_point(a.src, f, p, x);
        }
      }
      const int D = a.src.uv_dims;
      if (part == 0) {
        for (int d = 0; d < D; ++d) PE[d * TMP + m] = x[d];
        for (int c = E; c < e_chunks * KC; ++c) PE[c * TMP + m] = 0.f;
      }
      for (int k = part; k < kMultires; k += 4) {
        const float fr = (float)(1 << k);          // 2**linspace(0,9,10)
        for (int d = 0; d < D; ++d) {
          float sn, cs;
          sincosf(__fmul_rn(x[d], fr), &sn, &cs);
          PE[(D + (2 * k) * D + d) * TMP + m] = sn;
          PE[(D + (2 * k + 1) * D + d) * TMP + m] = cs;
        }
      }
      if (ROWLAT) {
        for (int k = part; k < kLatent; k += 4)
          LAT[k * TMP + m] = (p < Pf) ? a.rows[p * (a.src.uv_dims + kLatent) + a.src.uv_dims + k] : 0.f;
      } else {
        const float* fb = a.frame_bias + (size_t)f * 4 * 256;
        bias0[tid] = fb[tid];
        biasS[tid] = fb[256 + tid];
      }
    }
    __syncthreads();

    // ---- net = fc_uv(e) + fc_audio(a) + fc_time(t)            (tf_nerf.py:252-258)  -> X, no activation
    zero_acc(acc);
    gemm_seg(acc, PE, e_chunks, ps, a, wst, full, m0, tn);
    if (ROWLAT) gemm_seg(acc, LAT, 4, ps, a, wst, full, m0, tn);
    store_acc(acc, X, bias0, false, m0, tn, grow(0), rows_valid);
    __syncthreads();
    // ---- pts_linears 0..4 + ReLU                              (tf_nerf.py:265-267)
    float* in = X;
    float* out = Y;
    for (int l = 0; l < 5; ++l) {
      zero_acc(acc);
      gemm_seg(acc, in, 16, ps, a, wst, full, m0, tn);
      store_acc(acc, out, Fp + F_PTS_B + l * 256, true, m0, tn, grow(1 + l), rows_valid);
      __syncthreads();
      float* t = in; in = out; out = t;
    }
    // h (layer-4 output) is in Y, X is free
    // ---- h_skip = fc_uv_skip(e) + fc_audio_skip(a) + fc_time_skip(t)   (tf_nerf.py:268-276) -> X
    zero_acc(acc);
    gemm_seg(acc, PE, e_chunks, ps, a, wst, full, m0, tn);
    if (ROWLAT) gemm_seg(acc, LAT, 4, ps, a, wst, full, m0, tn);
    store_acc(acc, X, biasS, false, m0, tn, grow(6), rows_valid);
    __syncthreads();
    // ---- pts_linears.5 on cat([h_skip, h])                    (tf_nerf.py:281, :170-172) -> X
    zero_acc(acc);
    gemm_seg(acc, X, 16, ps, a, wst, full, m0, tn);
    gemm_seg(acc, Y, 16, ps, a, wst, full, m0, tn);
    store_acc(acc, X, Fp + F_PTS_B + 5 * 256, true, m0, tn, grow(7), rows_valid);
    __syncthreads();
    // ---- pts_linears 6, 7
    zero_acc(acc);
    gemm_seg(acc, X, 16, ps, a, wst, full, m0, tn);
    store_acc(acc, Y, Fp + F_PTS_B + 6 * 256, true, m0, tn, grow(8), rows_valid);
    __syncthreads();
    zero_acc(acc);
    gemm_seg(acc, Y, 16, ps, a, wst, full, m0, tn);
    store_acc(acc, X, Fp + F_PTS_B + 7 * 256, true, m0, tn, grow(9), rows_valid);
    __syncthreads();
    // ---- output_linear (raw, no activation)                   (tf_nerf.py:283)
    {
      const int m = tid & 63, n = tid >> 6;
      const long long p = p_base + m;
      float o = 0.f;
      if (n < a.out_ch) {
        const float* w = Fp + F_OUT_W + n * 256;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (int k = 0; k < 256; k += 4) {
          s0 = fmaf(X[(k + 0) * TMP + m], __ldg(w + k + 0), s0);
          s1 = fmaf(X[(k + 1) * TMP + m], __ldg(w + k + 1), s1);
          s2 = fmaf(X[(k + 2) * TMP + m], __ldg(w + k + 2), s2);
          s3 = fmaf(X[(k + 3) * TMP + m], __ldg(w + k + 3), s3);
        }
        o = ((s0 + s1) + (s2 + s3)) + Fp[F_OUT_B + n];
      }
      if (!ROWLAT && (a.fix_rgb || a.patch_raw)) {
        // listed rays (near-zero last-sample density under tensor-core rounding): this exact (r,g,b,sigma) of the
        // LAST sample either replaces the raw entry (unfused path) or finishes the pixel from the (T_last, acc) pair
        // the fused compositing epilogue left in `carry`  (rendering.py:43-60 with delta_last = 1e10)
        if (p < Pf) {
          const long long gray = (long long)f * a.src.R + a.src.list_rays[(long long)f * a.src.R + p];
          if (a.patch_raw) {
            if (n < a.out_ch) a.patch_raw[(gray * a.src.S + (a.src.S - 1)) * 4 + n] = o;
          } else {
            PE[n * TMP + m] = o;                       // PE is free after the skip GEMM
          }
        }
        if (a.fix_rgb) {
          __syncthreads();
          if (n == 0 && p < Pf) {
            const int ray = a.src.list_rays[(long long)f * a.src.R + p];
            const long long gray = (long long)f * a.src.R + ray;
            const float4 c = a.fix_carry[gray];
            const float nrm = ray_norm(a.src.rays_d + (a.src.rays_shared ? (long long)ray : gray) * 3);
            const float w = c.x * alpha_of(PE[3 * TMP + m], __fmul_rn(1e10f, nrm));
            a.fix_rgb[gray * 3 + 0] = fmaf(w, sigmoidf_acc(PE[0 * TMP + m]), c.y);
            a.fix_rgb[gray * 3 + 1] = fmaf(w, sigmoidf_acc(PE[1 * TMP + m]), c.z);
            a.fix_rgb[gray * 3 + 2] = fmaf(w, sigmoidf_acc(PE[2 * TMP + m]), c.w);
          }
        }
      } else if (n < a.out_ch && p < Pf) {
        a.out[((long long)f * a.src.P + p) * a.out_ch + n] = o;
      }
    }
    __syncthreads();
  }
}

static Fp32Program build_program(const Layout& L, int uv_dims, bool rowlat) {
  Fp32Program pr;
  int n = 0;
  const int fp = (int)(L.off_fp32 / 4), cc = (int)(L.off_const / 4);
  const int E = pe_dim(uv_dims), ech = (E + KC - 1) / KC;
  auto add = [&](int base, int nch) { for (int j = 0; j < nch; ++j) pr.off[n++] = base + j * CHUNK_FLOATS; };
  add(fp + F_UV_WT, ech);
  if (rowlat) add(cc + C_FCA_WT, 4);
  for (int l = 0; l < 5; ++l) add(fp + f_pts_off(l), 16);
  add(fp + F_UVS_WT, ech);
  if (rowlat) add(cc + C_FCAS_WT, 4);
  add(fp + f_pts_off(5), 32);
  add(fp + f_pts_off(6), 16);
  add(fp + f_pts_off(7), 16);
  pr.n_chunks = n;
  return pr;
}

static size_t fp32_smem_bytes(bool rowlat) {
  return sizeof(float) * (size_t)(2 * 256 * TMP + 64 * TMP + (rowlat ? 64 * TMP : 0) + NS * CHUNK_FLOATS);
}

int launch_mlp_fp32(const void* blob, const PointSrc& src, int n_frames, const float* frame_bias, float* out,
                    int out_ch, cudaStream_t st, const float4* fix_carry, float* fix_rgb, float* patch_raw) {
  Fp32Args a{};
  a.fix_carry = fix_carry;
  a.fix_rgb = fix_rgb;
  a.patch_raw = patch_raw;
  if ((fix_rgb || patch_raw) && (src.mode != S2L_PTS_RAYS || !src.list_rays || !src.tile_start || src.Sc != 1 || out_ch != 4)) {
    set_error("mlp_fp32: the last-sample re-evaluation needs a ray-list launch with Sc = 1 on the out_ch = 4 model");
    return 2;
  }
  a.blob = reinterpret_cast<const uint8_t*>(blob);
  a.L = blob_layout();
  a.src = src;
  a.frame_bias = frame_bias;
  a.out = out;
  a.out_ch = out_ch;
  a.n_frames = n_frames;
  a.tiles_per_frame = (src.P + TM - 1) / TM;
  a.prog = build_program(a.L, src.uv_dims, false);
  const long long n_tiles = a.tiles_per_frame * n_frames;
  if (n_tiles == 0) return 0;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = fp32_smem_bytes(false);
  static bool attr_set_dev[64] = {};   // cudaFuncSetAttribute is per device
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& attr_set = attr_set_dev[cur_dev & 63];
  if (!attr_set) {
    if (cudaFuncSetAttribute(mlp_fp32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("mlp_fp32: cannot opt in to %zu B of shared memory: %s", smem, cudaGetErrorString(cudaGetLastError()));
      return 6;
    }
    attr_set = true;
  }
  const unsigned grid = (unsigned)(n_tiles < sms ? n_tiles : sms);
  mlp_fp32_kernel<false><<<grid, 256, smem, st>>>(a);
  return check_launch("mlp_fp32_kernel") ? 0 : 5;
}

int launch_mlp_fp32_rows(const void* blob, const float* x, long long n_rows, long long time_idx, int has_time,
                         float* out, float* save, int uv_dims, int out_ch, cudaStream_t st, const Gate* gate,
                         const long long* time_idx_dev) {
  Fp32Args a{};
  if (gate) a.gate = *gate;
  a.time_idx_dev = time_idx_dev;
  a.blob = reinterpret_cast<const uint8_t*>(blob);
  a.L = blob_layout();
  a.src = PointSrc{};
  a.src.mode = S2L_PTS_EXPLICIT;
  a.src.uv_dims = uv_dims;
  a.src.P = n_rows;
  a.rows = x;
  a.save = save;
  a.time_idx = time_idx;
  a.has_time = has_time;
  a.out = out;
  a.out_ch = out_ch;
  a.n_frames = 1;
  a.tiles_per_frame = (n_rows + TM - 1) / TM;
  a.prog = build_program(a.L, uv_dims, true);
  if (n_rows == 0) return 0;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = fp32_smem_bytes(true);
  static bool attr_set_dev[64] = {};   // cudaFuncSetAttribute is per device
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& attr_set = attr_set_dev[cur_dev & 63];
  if (!attr_set) {
    if (cudaFuncSetAttribute(mlp_fp32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("mlp_fp32_rows: cannot opt in to %zu B of shared memory: %s", smem, cudaGetErrorString(cudaGetLastError()));
      return 6;
    }
    attr_set = true;
  }
  const unsigned grid = (unsigned)(a.tiles_per_frame < sms ? a.tiles_per_frame : sms);
  mlp_fp32_kernel<true><<<grid, 256, smem, st>>>(a);
  return check_launch("mlp_fp32_kernel<rows>") ? 0 : 5;
}

}  // namespace s2l
