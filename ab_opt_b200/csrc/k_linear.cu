// Small node-feature layers of the hot path on CUDA cores (FP32 FFMA, cp.async pipelined); the big GEMMs are in k_tc.cu.
//   mixer_kernel  : EpsilonNet input mixer + R = exp(v_t)           dpm_full.py:86-89
//   heads_kernel  : eps_crd / eps_rot / eps_seq / pRMSD heads + SO(3) update    dpm_full.py:92-110
#include <cstdlib>
#include "rowtile.cuh"
#include "params.cuh"
#include "kernels.h"

namespace abopt {

// ------------------------------------------------------------------------------------------ mixer
// rows / count (optional, inside the sampling loop): only the listed rows are evaluated -- everything the mixer reads of a context
// residue is a loop invariant there (res_feat, its sequence, its frame), so after the first step only the generated rows change.
// Results go to their place in x_out / x_lo_out / Rbuf / p_norm and, compact (row k of the list), to x_c / x_c_lo.
// R = rows per warp (rowtile.cuh): a CTA evaluates 8R rows.
template <int R>
__global__ void __launch_bounds__(RT_THREADS, 2)
mixer_kernel(int M, const float* __restrict__ res_feat, const long long* __restrict__ s_t,
             const float* __restrict__ v_t, EpsW w, float* __restrict__ x_out, float* __restrict__ Rbuf,
             const float* __restrict__ p_ang, float* __restrict__ p_norm, float mean0, float mean1, float mean2, float scale,
             float* __restrict__ x_lo_out, const int* __restrict__ rows, const int* __restrict__ count,
             float* __restrict__ x_c, float* __restrict__ x_c_lo) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RowTileSmem& s = *reinterpret_cast<RowTileSmem*>(smem_raw);
  constexpr int ROWS = 8 * R;
  float* act = reinterpret_cast<float*>(smem_raw + sizeof(RowTileSmem));      // [ROWS][RT_ACT_LD]
  __shared__ int aa[ROWS];
  __shared__ int ridx[ROWS];            // residue row of tile row r, -1 = none
  const int row0 = blockIdx.x * ROWS;
  const int nrows = rows ? count[0] : M;
  if (row0 >= nrows) return;
  if (threadIdx.x < ROWS) {
    const int k = row0 + threadIdx.x;
    const int r = (k < nrows) ? (rows ? rows[k] : k) : -1;
    ridx[threadIdx.x] = r;
    long long a = (r >= 0) ? s_t[r] : 0;
    aa[threadIdx.x] = (int)(a < 0 ? 0 : (a > 24 ? 24 : a));
    if (r >= 0 && p_ang != nullptr) {     // FullDPM._normalize_position (dpm_full.py:148-150)
      p_norm[(size_t)r * 3 + 0] = __fdiv_rn(__fadd_rn(p_ang[(size_t)r * 3 + 0], -mean0), scale);
      p_norm[(size_t)r * 3 + 1] = __fdiv_rn(__fadd_rn(p_ang[(size_t)r * 3 + 1], -mean1), scale);
      p_norm[(size_t)r * 3 + 2] = __fdiv_rn(__fadd_rn(p_ang[(size_t)r * 3 + 2], -mean2), scale);
    }
    if (r >= 0 && Rbuf != nullptr) {
      const Mat3 R = so3_exp(v_t[r * 3 + 0], v_t[r * 3 + 1], v_t[r * 3 + 2]);
#pragma unroll
      for (int i = 0; i < 9; ++i) Rbuf[(size_t)r * 9 + i] = R.m[i];
    }
  }
  __syncthreads();
  float acc[R][4];
  rt_zero(acc);
  // layer 0: [res_feat | embedding(s_t)] (K = 256) -> 128, ReLU
  rt_gemm_globalA(acc, s, [&](int r, int k) -> const float* {
    if (ridx[r] < 0) return nullptr;
    return (k < F) ? res_feat + (size_t)ridx[r] * F + k : w.emb + (size_t)aa[r] * F + (k - F);
  }, w.Wm0_t, 2 * F);
  rt_add_bias(acc, w.bm0);
  rt_relu(acc);
  rt_store_act(acc, act, RT_ACT_LD);
  __syncthreads();
  rt_zero(acc);
  rt_gemm_smemA(acc, s, act, RT_ACT_LD, w.Wm2_t, F);
  rt_add_bias(acc, w.bm2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int row = ridx[warp * R + r];
    if (row >= 0) {
      const float4 hi = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      const float4 lo = make_float4(tf32_lo(acc[r][0]), tf32_lo(acc[r][1]), tf32_lo(acc[r][2]), tf32_lo(acc[r][3]));
      *reinterpret_cast<float4*>(x_out + (size_t)row * F + lane * 4) = hi;
      if (x_lo_out != nullptr) *reinterpret_cast<float4*>(x_lo_out + (size_t)row * F + lane * 4) = lo;
      if (x_c != nullptr) {
        const size_t kc = (size_t)(row0 + warp * R + r) * F + lane * 4;
        *reinterpret_cast<float4*>(x_c + kc) = hi;
        *reinterpret_cast<float4*>(x_c_lo + kc) = lo;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ heads
constexpr int HD_OUT_LD = 68;      // staged head outputs per row: crd 0-2 | rot 3-5 | seq 6-25 | prmsd 26-65

// One 3-layer head over the 8R-row tile.  `in_act` holds the 128 node features; the 3 time-embedding
// inputs enter as a per-row rank-3 correction `ext[r][q]` (q = 0..2) times W0_ext.
template <int R>
__device__ __forceinline__ void run_head(float (&acc)[R][4], RowTileSmem& s, const float* in_act, float* hid,
                                         const HeadW& hw, const float (&ext)[R][3]) {
  const int lane = threadIdx.x & 31;
  rt_zero(acc);
  rt_gemm_smemA(acc, s, in_act, RT_ACT_LD, hw.W0_t, F);
  rt_add_bias(acc, hw.b0);
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const float4 we = *reinterpret_cast<const float4*>(hw.W0_ext + q * F + lane * 4);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      acc[r][0] = fmaf(ext[r][q], we.x, acc[r][0]); acc[r][1] = fmaf(ext[r][q], we.y, acc[r][1]);
      acc[r][2] = fmaf(ext[r][q], we.z, acc[r][2]); acc[r][3] = fmaf(ext[r][q], we.w, acc[r][3]);
    }
  }
  rt_relu(acc);
  __syncthreads();                       // everyone is done reading `hid` from a previous head
  rt_store_act(acc, hid, RT_ACT_LD);
  __syncthreads();
  rt_zero(acc);
  rt_gemm_smemA(acc, s, hid, RT_ACT_LD, hw.W2_t, F);
  rt_add_bias(acc, hw.b2);
  rt_relu(acc);
  __syncthreads();
  rt_store_act(acc, hid, RT_ACT_LD);
  __syncthreads();
  rt_zero(acc);
  rt_gemm_smemA(acc, s, hid, RT_ACT_LD, hw.W4_t, F);
  rt_add_bias(acc, hw.b4);
}

// copy output columns [0, n) of the tile into the staging buffer at column offset `off`
template <int R>
__device__ __forceinline__ void stage_out(const float (&acc)[R][4], float* outs, int off, int n) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int col = lane * 4 + c;
    if (col < n)
#pragma unroll
      for (int r = 0; r < R; ++r) outs[(warp * R + r) * HD_OUT_LD + off + col] = acc[r][c];
  }
}

template <int R>
__global__ void __launch_bounds__(RT_THREADS, 1)
heads_kernel(int M, int L, const float* __restrict__ x, const float* __restrict__ beta, int beta_stride,
             const float* __restrict__ Rbuf, const float* __restrict__ v_t, const uint8_t* __restrict__ mask_gen,
             EpsW w, float* __restrict__ v_next, float* __restrict__ R_next, float* __restrict__ eps_pos,
             float* __restrict__ c_den, float* __restrict__ prmsd_rows, const int* __restrict__ rows, const int* __restrict__ count,
             int head_lo, int x_compact) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // focus mode: only the `count[0]` rows listed in rows[] are evaluated.  x is either compact (x_compact: row k of x is
  // residue rows[k], the output of a focused last block) or the full tensor; every other tensor is indexed by residue row.
  constexpr int ROWS = 8 * R;
  if (count) { const int g = count[0]; M = g < M ? g : M; }
  if ((int)blockIdx.x * ROWS >= M) return;
  RowTileSmem& s = *reinterpret_cast<RowTileSmem*>(smem_raw);
  float* xin = reinterpret_cast<float*>(smem_raw + sizeof(RowTileSmem));      // [ROWS][RT_ACT_LD]
  float* hid = xin + ROWS * RT_ACT_LD;                                        // [ROWS][RT_ACT_LD]
  float* outs = hid + ROWS * RT_ACT_LD;                                       // [ROWS][HD_OUT_LD]
  const int row0 = blockIdx.x * ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // resident input tile + per-row time embedding (dpm_full.py:92-93)
  float ext[R][3];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int row = row0 + warp * R + r;
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
    float b = 0.f;
    if (row < M) {
      const int rr = rows ? rows[row] : row;
      xv = *reinterpret_cast<const float4*>(x + (size_t)(x_compact ? row : rr) * F + lane * 4);
      b = beta[(size_t)(rr / L) * beta_stride];
    }
    *reinterpret_cast<float4*>(xin + (warp * R + r) * RT_ACT_LD + lane * 4) = xv;
    ext[r][0] = b; ext[r][1] = sinf(b); ext[r][2] = cosf(b);
  }
  __syncthreads();

  // blockIdx.y selects the head (0 eps_crd, 1 eps_rot, 2 eps_seq, 3 pRMSD): the four heads are independent down to their
  // per-residue epilogues, so they run as separate CTAs (4x the parallelism of one CTA walking all heads; in focus mode
  // only a few row tiles exist)
  const int head = head_lo + blockIdx.y;
  float acc[R][4];
  if (head == 0) { run_head(acc, s, xin, hid, w.crd, ext); stage_out(acc, outs, 0, 3); }
  else if (head == 1) { run_head(acc, s, xin, hid, w.rot, ext); stage_out(acc, outs, 3, 3); }
  else if (head == 2) { run_head(acc, s, xin, hid, w.seq, ext); stage_out(acc, outs, 6, NAA); }
  else {
    // PerResiduePredictor (common/nn.py:164-188): LayerNorm over the 131 inputs, then 3 linears.
    float g[R][4], gext[R][3];
    const float4 lg = *reinterpret_cast<const float4*>(w.prm_ln_g + lane * 4);
    const float4 lb = *reinterpret_cast<const float4*>(w.prm_ln_b + lane * 4);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float4 xv = *reinterpret_cast<const float4*>(xin + (warp * R + r) * RT_ACT_LD + lane * 4);
      const float sum = warp_sum(xv.x + xv.y + xv.z + xv.w) + (ext[r][0] + ext[r][1] + ext[r][2]);
      const float mean = sum * (1.f / 131.f);
      const float d0 = xv.x - mean, d1 = xv.y - mean, d2 = xv.z - mean, d3 = xv.w - mean;
      const float e0 = ext[r][0] - mean, e1 = ext[r][1] - mean, e2 = ext[r][2] - mean;
      const float var = (warp_sum(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3) + (e0 * e0 + e1 * e1 + e2 * e2)) * (1.f / 131.f);
      const float sd = sqrtf(var + 1e-10f);
      g[r][0] = d0 / sd * lg.x + lb.x; g[r][1] = d1 / sd * lg.y + lb.y;
      g[r][2] = d2 / sd * lg.z + lb.z; g[r][3] = d3 / sd * lg.w + lb.w;
      gext[r][0] = e0 / sd * w.prm_ln_g[128] + w.prm_ln_b[128];
      gext[r][1] = e1 / sd * w.prm_ln_g[129] + w.prm_ln_b[129];
      gext[r][2] = e2 / sd * w.prm_ln_g[130] + w.prm_ln_b[130];
    }
    __syncthreads();                                   // every warp has read xin
    rt_store_act(g, xin, RT_ACT_LD);
    __syncthreads();
    run_head(acc, s, xin, hid, w.prm, gext);
    stage_out(acc, outs, 26, w.prmsd_bins);
  }
  __syncthreads();

  // per-residue epilogue of this head
  if (threadIdx.x < ROWS) {
    const int krow = row0 + threadIdx.x;
    if (krow < M) {
      const int row = rows ? rows[krow] : krow;
      const float* o = outs + threadIdx.x * HD_OUT_LD;
      const bool gen = mask_gen[row] != 0;
      if (head == 0) {
        // eps_pos = R eps_crd, zero outside the generated region (dpm_full.py:96-98)
        Mat3 R;
#pragma unroll
        for (int i = 0; i < 9; ++i) R.m[i] = Rbuf[(size_t)row * 9 + i];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float e = R.m[i * 3 + 0] * o[0] + R.m[i * 3 + 1] * o[1] + R.m[i * 3 + 2] * o[2] + 0.f;
          eps_pos[(size_t)row * 3 + i] = gen ? e : 0.f;
        }
      } else if (head == 1) {
        // R_next = R U(eps_rot), v_next = log(R_next) on generated residues (dpm_full.py:101-105)
        Mat3 R;
#pragma unroll
        for (int i = 0; i < 9; ++i) R.m[i] = Rbuf[(size_t)row * 9 + i];
        const Mat3 U = quat1ijk_to_rot(o[3], o[4], o[5]);
        const Mat3 Rn = matmul3(R, U);
        if (R_next != nullptr)
#pragma unroll
          for (int i = 0; i < 9; ++i) R_next[(size_t)row * 9 + i] = Rn.m[i];
        float vx, vy, vz;
        so3_log(Rn, vx, vy, vz);
        v_next[(size_t)row * 3 + 0] = gen ? vx : v_t[(size_t)row * 3 + 0];
        v_next[(size_t)row * 3 + 1] = gen ? vy : v_t[(size_t)row * 3 + 1];
        v_next[(size_t)row * 3 + 2] = gen ? vz : v_t[(size_t)row * 3 + 2];
      } else if (head == 2) {
        // softmax over 20 classes (dpm_full.py:61,108)
        float mx = o[6];
#pragma unroll
        for (int k = 1; k < NAA; ++k) mx = fmaxf(mx, o[6 + k]);
        float e[NAA], sum = 0.f;
#pragma unroll
        for (int k = 0; k < NAA; ++k) { e[k] = expf(o[6 + k] - mx); sum += e[k]; }
#pragma unroll
        for (int k = 0; k < NAA; ++k) c_den[(size_t)row * NAA + k] = e[k] / sum;
      } else {
        for (int k = 0; k < w.prmsd_bins; ++k) prmsd_rows[(size_t)row * w.prmsd_bins + k] = o[26 + k];
      }
    }
  }
}

// prmsd_logits[n][k] = mean over ALL L rows, padding included (dpm_full.py:110).  One CTA per complex,
// fixed summation order (deterministic).
__global__ void prmsd_mean_kernel(int L, int bins, const float* __restrict__ prmsd_rows, float* __restrict__ out) {
  const int n = blockIdx.x, k = threadIdx.x;
  if (k >= bins) return;
  float sum = 0.f;
  for (int l = 0; l < L; ++l) sum += prmsd_rows[((size_t)n * L + l) * bins + k];
  out[(size_t)n * bins + k] = sum / (float)L;
}

// ------------------------------------------------------------------------------------------ focus (last-layer row restriction)
// Inside the sampling loop of a model WITHOUT the pRMSD head, the output of the last GABlock is consumed only by the three
// heads, and their outputs only on generated residues (dpm_full.py:98,105; transition.py:99,158,176).  So the last block
// needs its query side (logits rows, pair / node / point aggregation, tail) only for the generated rows.  focus_build_kernel
// turns mask_generate into: the compact row list, its inverse, and per complex the 128-row query windows that cover them.
__global__ void __launch_bounds__(256)
focus_build_kernel(int N, int L, const uint8_t* __restrict__ mask_gen, int* __restrict__ cidx, int* __restrict__ rows,
                   int2* __restrict__ windows, int* __restrict__ count, int* __restrict__ scratch) {
  int* nrow = scratch;            // [N] generated rows per complex -> exclusive prefix
  int* nwin = scratch + N;        // [N] windows per complex -> exclusive prefix
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    int g = 0, wct = 0, cover = -1;
    for (int i = 0; i < L; ++i)
      if (mask_gen[(size_t)n * L + i]) {
        ++g;
        if (i >= cover) { ++wct; cover = (i & ~7) + 128; }
      }
    nrow[n] = g; nwin[n] = wct;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int a = 0, b = 0;
    for (int n = 0; n < N; ++n) { const int g = nrow[n], wv = nwin[n]; nrow[n] = a; nwin[n] = b; a += g; b += wv; }
    count[0] = a; count[1] = b;
  }
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    int k = nrow[n], wk = nwin[n], cover = -1;
    for (int i = 0; i < L; ++i) {
      const size_t r = (size_t)n * L + i;
      if (mask_gen[r]) {
        cidx[r] = k; rows[k] = (int)r; ++k;
        if (i >= cover) { windows[wk++] = make_int2(n, i & ~7); cover = (i & ~7) + 128; }
      } else cidx[r] = -1;
    }
  }
}
// x_c[k] = x[rows[k]], mask_c[k] = mask[rows[k]]: the residual input and the residue mask of the compact tail
__global__ void focus_gather_kernel(const int* __restrict__ rows, const int* __restrict__ count, const float* __restrict__ x,
                                    const uint8_t* __restrict__ mask, float* __restrict__ x_c, uint8_t* __restrict__ mask_c) {
  const int g = count[0];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, nwarp = (gridDim.x * blockDim.x) >> 5;
  for (int k = warp; k < g; k += nwarp) {
    const int r = rows[k];
    reinterpret_cast<float4*>(x_c + (size_t)k * F)[lane] = reinterpret_cast<const float4*>(x + (size_t)r * F)[lane];
    if (lane == 0) mask_c[k] = mask[r];
  }
}
void launch_focus_build(int N, int L, const uint8_t* mask_gen, int* cidx, int* rows, int2* windows, int* count, int* scratch,
                        cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  focus_build_kernel<<<1, 256, 0, st>>>(N, L, mask_gen, cidx, rows, windows, count, scratch);
}
void launch_focus_gather(const int* rows, const int* count, const float* x, const uint8_t* mask, float* x_c, uint8_t* mask_c,
                         cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  focus_gather_kernel<<<64, 256, 0, st>>>(rows, count, x, mask, x_c, mask_c);
}

// ------------------------------------------------------------------------------------------ launchers
// Rows per warp (R) of the two row-tile kernels.  Whole-batch launches use R = 8 (64-row tiles, the best FMA : shared-memory
// ratio).  The launches of the sampling loop that walk a row list (the generated residues: 1 024 rows at C2) are bound by the
// latency of ONE tile, not by throughput -- 16 tiles of 64 rows leave 132 SMs idle -- so they use 16-row tiles on more SMs.
// The results do not depend on R (rowtile.cuh; scripts/ub/rpw_ab.py checks a whole C2 sample bit for bit).  Measured on B200
// (profiles/r02_rpw_ab.jsonl, per C2 sample of 100 steps): mixer 3.80 / 2.79 / 2.39 ms and heads 5.03 / 4.00 / 3.53 ms at
// R = 8 / 4 / 2.  ABOPT_RPW_MIXER / ABOPT_RPW_HEADS (2, 4 or 8; read per call) override the list-launch defaults for A/B runs.
constexpr int MIXER_LIST_RPW = 2, HEADS_LIST_RPW = 2;

static int list_rpw(const char* env_name, int dflt) {
  const char* e = getenv(env_name);
  if (e == nullptr || e[0] == '\0') return dflt;
  const int v = atoi(e);
  return (v == 2 || v == 4 || v == 8) ? v : dflt;
}

template <int R> static size_t mixer_smem() { return sizeof(RowTileSmem) + 8 * R * RT_ACT_LD * sizeof(float); }
template <int R> static size_t heads_smem() { return sizeof(RowTileSmem) + (2 * 8 * R * RT_ACT_LD + 8 * R * HD_OUT_LD) * sizeof(float); }

template <int R> static cudaError_t linear_init_r() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(mixer_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mixer_smem<R>())) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(heads_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)heads_smem<R>())) != cudaSuccess) return e;
  return cudaSuccess;
}
cudaError_t linear_kernels_init() {
  cudaError_t e;
  if ((e = linear_init_r<8>()) != cudaSuccess) return e;
  if ((e = linear_init_r<4>()) != cudaSuccess) return e;
  if ((e = linear_init_r<2>()) != cudaSuccess) return e;
  return cudaSuccess;
}

template <int R>
static void mixer_launch_r(int M, const float* res_feat, const long long* s_t, const float* v_t, const EpsW& w, float* x_out,
                           float* Rbuf, const float* p_ang, float* p_norm, const float* mean, float scale, float* x_lo_out,
                           cudaStream_t st, const int* rows, const int* count, float* x_c, float* x_c_lo) {
  mixer_kernel<R><<<(M + 8 * R - 1) / (8 * R), RT_THREADS, mixer_smem<R>(), st>>>(M, res_feat, s_t, v_t, w, x_out, Rbuf, p_ang, p_norm,
                                                                             mean[0], mean[1], mean[2], scale, x_lo_out, rows, count,
                                                                             x_c, x_c_lo);
}
void launch_mixer(int M, const float* res_feat, const long long* s_t, const float* v_t, const EpsW& w,
                  float* x_out, float* Rbuf, const float* p_ang, float* p_norm, const float* mean, float scale,
                  float* x_lo_out, cudaStream_t st, const int* rows, const int* count, float* x_c, float* x_c_lo) {
  ProfScope prof__(KK_MIXER, st);
  const int R = rows ? list_rpw("ABOPT_RPW_MIXER", MIXER_LIST_RPW) : 8;
  if (R == 2) mixer_launch_r<2>(M, res_feat, s_t, v_t, w, x_out, Rbuf, p_ang, p_norm, mean, scale, x_lo_out, st, rows, count, x_c, x_c_lo);
  else if (R == 4) mixer_launch_r<4>(M, res_feat, s_t, v_t, w, x_out, Rbuf, p_ang, p_norm, mean, scale, x_lo_out, st, rows, count, x_c, x_c_lo);
  else mixer_launch_r<8>(M, res_feat, s_t, v_t, w, x_out, Rbuf, p_ang, p_norm, mean, scale, x_lo_out, st, rows, count, x_c, x_c_lo);
}

// one launch of `nheads` heads starting at head_lo over the row tiles of M rows (or of the row list)
template <int R>
static void heads_launch_r(int nheads, int head_lo, int M, int L, const float* x, const float* beta, int beta_stride, const float* Rbuf,
                           const float* v_t, const uint8_t* mask_gen, const EpsW& w, float* v_next, float* R_next, float* eps_pos,
                           float* c_den, float* prmsd_rows, cudaStream_t st, const int* rows, const int* count, int x_compact) {
  heads_kernel<R><<<dim3((M + 8 * R - 1) / (8 * R), nheads), RT_THREADS, heads_smem<R>(), st>>>(
      M, L, x, beta, beta_stride, Rbuf, v_t, mask_gen, w, v_next, R_next, eps_pos, c_den, prmsd_rows, rows, count, head_lo, x_compact);
}
void launch_heads(int M, int L, const float* x, const float* beta, int beta_stride, const float* Rbuf, const float* v_t,
                  const uint8_t* mask_gen, const EpsW& w, float* v_next, float* R_next, float* eps_pos, float* c_den,
                  float* prmsd_rows, float* prmsd_logits, cudaStream_t st, const int* rows, const int* count, bool x_compact) {
  ProfScope prof__(KK_HEADS, st);
  const int R = rows ? list_rpw("ABOPT_RPW_HEADS", HEADS_LIST_RPW) : 8;
  auto go = [&](int r, int nheads, int head_lo, const int* rws, const int* cnt, int xc) {
    if (r == 2) heads_launch_r<2>(nheads, head_lo, M, L, x, beta, beta_stride, Rbuf, v_t, mask_gen, w, v_next, R_next, eps_pos, c_den, prmsd_rows, st, rws, cnt, xc);
    else if (r == 4) heads_launch_r<4>(nheads, head_lo, M, L, x, beta, beta_stride, Rbuf, v_t, mask_gen, w, v_next, R_next, eps_pos, c_den, prmsd_rows, st, rws, cnt, xc);
    else heads_launch_r<8>(nheads, head_lo, M, L, x, beta, beta_stride, Rbuf, v_t, mask_gen, w, v_next, R_next, eps_pos, c_den, prmsd_rows, st, rws, cnt, xc);
  };
  if (rows && w.has_prmsd) {
    // the crd / rot / seq heads on the listed (generated) rows only, the pRMSD head on every row (its logits are averaged
    // over all L rows, dpm_full.py:110)
    go(R, 3, 0, rows, count, x_compact ? 1 : 0);
    count_launch();
    go(8, 1, 3, nullptr, nullptr, 0);
  } else {
    go(R, w.has_prmsd ? 4 : 3, 0, rows, count, x_compact ? 1 : 0);
  }
  if (w.has_prmsd && prmsd_logits != nullptr) {
    prmsd_mean_kernel<<<M / L, 64, 0, st>>>(L, w.prmsd_bins, prmsd_rows, prmsd_logits);
    count_launch();
  }
}

}  // namespace abopt
