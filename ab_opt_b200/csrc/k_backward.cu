// Backward pass of the training step: `loss.backward()` of
// /root/reference/AbDock/train.py:104-113 through FullDPM.forward (modules/diffusion/dpm_full.py:156-234; AbDesign :138-190),
// EpsilonNet.forward (:70-112) and GABlock.forward (modules/encoders/ga.py:149-178), written out by hand.  The reference has no
// backward code of its own (torch autograd); the formulas implemented here are the ones of oracle/ipa_backward.py and
// oracle/epsnet_backward.py, which are pinned to the gradients the unmodified reference computes (tests/golden/train_backward.npz).
//
// Scheme: recompute-based.  The forward pass keeps only the block inputs x_l; for every block (last to first) alpha and the
// 1824-wide aggregate are recomputed by the forward tensor-core kernels, the plain projections by a GEMM, and the gradients flow
//   tail (LN2, MLP, LN1, out_transform)  ->  aggregate  ->  alpha  ->  softmax  ->  logits  ->  projections  ->  x, z, weights.
// Kernels of this file:
//   gemm_f32_kernel        generic strided fp32 GEMM on the CUDA cores (activation GEMMs NT / NN, weight gradients TN with a
//                          deterministic split over the long row dimension)
//   row kernels            LayerNorm backward, tail forward with saved activations, aggregate backward, column sums
//   pair_bwd_query_kernel  one CTA per query row (b, i): streams z[b,i,:,:] once, d alpha -> softmax backward -> d logits,
//                          d z (written or accumulated), d q / d query points, partial sums for d W_b and d spatial_coef
//   pair_bwd_key_kernel    one warp per (b, h, 32 keys): the transposed contractions d k, d key points, d v, d value points
// This first version favours exactness and simplicity over speed (fp32 FFMA everywhere, gradients reproducible run to run).
#include <cmath>
#include "kernels.h"
#define ABOPT_MAX_L_INTERNAL 512

namespace abopt {

// ------------------------------------------------------------------------------------------ generic fp32 GEMM
// C[i][j] (+)= sum_k a(i, k) b(k, j) (+ bias[j]),  a(i, k) = A[i sa_i + k sa_k] (optionally max(., 0)),  b(k, j) = B[k sb_k + j sb_j]
// (optionally max(., 0)).  64 x 64 output tile per CTA, 256 threads x (4 x 4), k-step 16.  splits > 1: grid.z slices the k
// range and writes partial tiles to `part` [split][M][N]; gemm_reduce_kernel adds them up in a fixed order.
struct GemmArgs {
  int M, N, K;
  const float* A; long long sa_i, sa_k;
  const float* B; long long sb_k, sb_j;
  float* C; int ldc;
  const float* bias;
  int accumulate, relu_a, relu_b;
  int k_per_split;
  float* part;
};

__global__ void __launch_bounds__(256) gemm_f32_kernel(const GemmArgs g) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  const int k_lo = blockIdx.z * g.k_per_split;
  const int k_hi = min(g.K, k_lo + g.k_per_split);
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  // loaders: thread -> (k, i) so that the unit-stride direction is contiguous across consecutive threads
  const bool a_k_fast = g.sa_k == 1, b_k_fast = g.sb_k == 1;
  for (int k0 = k_lo; k0 < k_hi; k0 += 16) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = threadIdx.x + e * 256;                 // 1024 elements of a 64 x 16 tile
      const int kk = a_k_fast ? (idx & 15) : (idx >> 6), ii = a_k_fast ? (idx >> 4) : (idx & 63);
      const int gi = i0 + ii, gk = k0 + kk;
      float v = (gi < g.M && gk < k_hi) ? g.A[(long long)gi * g.sa_i + (long long)gk * g.sa_k] : 0.f;
      if (g.relu_a) v = fmaxf(v, 0.f);
      As[kk][ii] = v;
      const int kb = b_k_fast ? (idx & 15) : (idx >> 6), jj = b_k_fast ? (idx >> 4) : (idx & 63);
      const int gj = j0 + jj, gkb = k0 + kb;
      float w = (gj < g.N && gkb < k_hi) ? g.B[(long long)gkb * g.sb_k + (long long)gj * g.sb_j] : 0.f;
      if (g.relu_b) w = fmaxf(w, 0.f);
      Bs[kb][jj] = w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a4[r], b4[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int gi = i0 + ty * 4 + r;
    if (gi >= g.M) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int gj = j0 + tx * 4 + c;
      if (gj >= g.N) continue;
      if (g.part) { g.part[((size_t)blockIdx.z * g.M + gi) * g.N + gj] = acc[r][c]; continue; }
      float v = acc[r][c] + (g.bias ? g.bias[gj] : 0.f);
      float* dst = g.C + (size_t)gi * g.ldc + gj;
      *dst = g.accumulate ? *dst + v : v;
    }
  }
}

__global__ void gemm_reduce_kernel(int MN, int N, int splits, const float* __restrict__ part, float* __restrict__ C, int ldc, int accumulate) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= MN) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[(size_t)z * MN + e];
  float* dst = C + (size_t)(e / N) * ldc + (e % N);
  *dst = accumulate ? *dst + s : s;
}

static void gemm(const GemmArgs& a0, int splits, cudaStream_t st) {
  GemmArgs a = a0;
  ProfScope prof__(KK_OTHER, st);
  a.k_per_split = splits > 1 ? ((a.K + splits - 1) / splits + 15) / 16 * 16 : a.K;
  if (splits <= 1) a.part = nullptr;
  dim3 grid((a.N + 63) / 64, (a.M + 63) / 64, splits > 1 ? splits : 1);
  gemm_f32_kernel<<<grid, 256, 0, st>>>(a);
  if (splits > 1) {
    const int MN = a.M * a.N;
    gemm_reduce_kernel<<<(MN + 255) / 256, 256, 0, st>>>(MN, a.N, splits, a.part, a.C, a.ldc, a.accumulate);
  }
}
// y[M][N] = x[M][K] W[N][K]^T (+ bias)           (nn.Linear forward)
void bwd_gemm_nt(int M, int N, int K, const float* x, int ldx, bool relu_x, const float* W, int ldw, const float* bias, float* y, int ldy,
                 cudaStream_t st) {
  gemm(GemmArgs{M, N, K, x, ldx, 1, W, 1, ldw, y, ldy, bias, 0, relu_x ? 1 : 0, 0, 0, nullptr}, 1, st);
}
// dx[M][K] (+)= g[M][N] W[N][K]                  (nn.Linear backward, input gradient)
void bwd_gemm_nn(int M, int N, int K, const float* g, int ldg, const float* W, int ldw, float* dx, int lddx, bool accumulate, cudaStream_t st) {
  gemm(GemmArgs{M, K, N, g, ldg, 1, W, ldw, 1, dx, lddx, nullptr, accumulate ? 1 : 0, 0, 0, 0, nullptr}, 1, st);
}
// dW[N][K] = g[M][N]^T x[M][K]                   (nn.Linear backward, weight gradient); scratch: splits * N * K floats
void bwd_wgrad(int M, int N, int K, const float* g, int ldg, const float* x, int ldx, bool relu_x, float* dW, float* scratch,
               size_t scratch_floats, cudaStream_t st) {
  int splits = (M + 511) / 512;
  if (splits > 64) splits = 64;
  while (splits > 1 && (size_t)splits * N * K > scratch_floats) --splits;
  gemm(GemmArgs{N, K, M, g, 1, ldg, x, ldx, 1, dW, K, nullptr, 0, 0, relu_x ? 1 : 0, 0, scratch}, splits, st);
}

// ------------------------------------------------------------------------------------------ column sums (bias / gamma / beta gradients)
// out[c] = sum_r A[r][c] (* B[r][c]); two deterministic stages
__global__ void colsum_stage1(int M, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, int rows_per,
                              float* __restrict__ part) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= K) return;
  const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += B ? A[(size_t)r * lda + c] * B[(size_t)r * ldb + c] : A[(size_t)r * lda + c];
  part[(size_t)blockIdx.y * K + c] = s;
}
__global__ void colsum_stage2(int K, int nparts, const float* __restrict__ part, float* __restrict__ out, float scale) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= K) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(size_t)p * K + c];
  out[c] = s * scale;
}
void bwd_colsum(int M, int K, const float* A, int lda, const float* B, int ldb, float* out, float* scratch, cudaStream_t st, float scale) {
  ProfScope prof__(KK_OTHER, st);
  const int rows_per = 256, nparts = (M + rows_per - 1) / rows_per;
  colsum_stage1<<<dim3((K + 127) / 128, nparts), 128, 0, st>>>(M, K, A, lda, B, ldb, rows_per, scratch);
  colsum_stage2<<<(K + 127) / 128, 128, 0, st>>>(K, nparts, scratch, out, scale);
}

// ------------------------------------------------------------------------------------------ row kernels (one warp per row of 128)
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// s = a + (mask ? b : 0) ; h = LayerNorm(s) (layers.py:146-155: biased variance, eps 1e-10 inside the square root)
__global__ void add_ln_fwd_kernel(int M, const float* __restrict__ a, const float* __restrict__ b, const uint8_t* __restrict__ mask,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ s_out,
                                  float* __restrict__ h_out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const bool mk = mask ? mask[row] != 0 : true;
  float4 av = *reinterpret_cast<const float4*>(a + (size_t)row * F + lane * 4);
  const float4 bv = *reinterpret_cast<const float4*>(b + (size_t)row * F + lane * 4);
  if (mk) { av.x += bv.x; av.y += bv.y; av.z += bv.z; av.w += bv.w; }
  if (s_out) *reinterpret_cast<float4*>(s_out + (size_t)row * F + lane * 4) = av;
  if (!h_out) return;
  const float mu = wsum(av.x + av.y + av.z + av.w) * (1.f / F);
  const float d0 = av.x - mu, d1 = av.y - mu, d2 = av.z - mu, d3 = av.w - mu;
  const float inv = 1.f / sqrtf(wsum(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3) * (1.f / F) + 1e-10f);
  const float4 gm = *reinterpret_cast<const float4*>(gamma + lane * 4), bt = *reinterpret_cast<const float4*>(beta + lane * 4);
  *reinterpret_cast<float4*>(h_out + (size_t)row * F + lane * 4) =
      make_float4(d0 * inv * gm.x + bt.x, d1 * inv * gm.y + bt.y, d2 * inv * gm.z + bt.z, d3 * inv * gm.w + bt.w);
}
// LayerNorm backward (oracle/ipa_backward.py:_layer_norm_backward): g (+ g2) -> ds; gxh = g * xhat (for d gamma), gsum = g (for d beta)
__global__ void ln_bwd_kernel(int M, const float* __restrict__ g, const float* __restrict__ g2, const float* __restrict__ s,
                              const float* __restrict__ gamma, float* __restrict__ ds, float* __restrict__ gxh, float* __restrict__ gsum) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  float4 gv = *reinterpret_cast<const float4*>(g + (size_t)row * F + lane * 4);
  if (g2) { const float4 t = *reinterpret_cast<const float4*>(g2 + (size_t)row * F + lane * 4); gv.x += t.x; gv.y += t.y; gv.z += t.z; gv.w += t.w; }
  const float4 sv = *reinterpret_cast<const float4*>(s + (size_t)row * F + lane * 4);
  const float mu = wsum(sv.x + sv.y + sv.z + sv.w) * (1.f / F);
  const float d0 = sv.x - mu, d1 = sv.y - mu, d2 = sv.z - mu, d3 = sv.w - mu;
  const float inv = 1.f / sqrtf(wsum(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3) * (1.f / F) + 1e-10f);
  const float x0 = d0 * inv, x1 = d1 * inv, x2 = d2 * inv, x3 = d3 * inv;
  const float4 gm = *reinterpret_cast<const float4*>(gamma + lane * 4);
  const float a0 = gv.x * gm.x, a1 = gv.y * gm.y, a2 = gv.z * gm.z, a3 = gv.w * gm.w;
  const float m1 = wsum(a0 + a1 + a2 + a3) * (1.f / F);
  const float m2 = wsum(a0 * x0 + a1 * x1 + a2 * x2 + a3 * x3) * (1.f / F);
  *reinterpret_cast<float4*>(ds + (size_t)row * F + lane * 4) =
      make_float4(inv * (a0 - m1 - x0 * m2), inv * (a1 - m1 - x1 * m2), inv * (a2 - m1 - x2 * m2), inv * (a3 - m1 - x3 * m2));
  *reinterpret_cast<float4*>(gxh + (size_t)row * F + lane * 4) = make_float4(gv.x * x0, gv.y * x1, gv.z * x2, gv.w * x3);
  if (gsum) *reinterpret_cast<float4*>(gsum + (size_t)row * F + lane * 4) = gv;
}
// g *= (a > 0)   (ReLU backward), n elements
__global__ void relu_bwd_kernel(size_t n, float* __restrict__ g, const float* __restrict__ a) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n && !(a[i] > 0.f)) g[i] = 0.f;
}
// out = a (+ b), rows masked to zero where mask == 0 (mask may be null)
__global__ void add_mask_kernel(int M, int K, const float* a, const float* __restrict__ b, const uint8_t* __restrict__ mask,
                                float* out) {      // (out may be a: in-place accumulation)
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)M * K) return;
  const int r = (int)(i / K);
  float v = a[i] + (b ? b[i] : 0.f);
  if (mask && mask[r] == 0) v = 0.f;
  out[i] = v;
}
void bwd_add_ln_fwd(int M, const float* a, const float* b, const uint8_t* mask, const float* gamma, const float* beta, float* s_out,
                    float* h_out, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  add_ln_fwd_kernel<<<(M + 7) / 8, 256, 0, st>>>(M, a, b, mask, gamma, beta, s_out, h_out);
}
void bwd_ln(int M, const float* g, const float* g2, const float* s, const float* gamma, float* ds, float* gxh, float* gsum, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  ln_bwd_kernel<<<(M + 7) / 8, 256, 0, st>>>(M, g, g2, s, gamma, ds, gxh, gsum);
}
void bwd_relu(size_t n, float* g, const float* a, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  relu_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, g, a);
}
void bwd_add_mask(int M, int K, const float* a, const float* b, const uint8_t* mask, float* out, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  add_mask_kernel<<<(unsigned)(((size_t)M * K + 255) / 256), 256, 0, st>>>(M, K, a, b, mask, out);
}

// ------------------------------------------------------------------------------------------ points: local -> global frame
// PG[row][kind * 288 + c] = R p + t for the query / key / value points of the plain projections P (M x 2016)   (geometry.py:72-91)
__global__ void points_global_kernel(int M, const float* __restrict__ Pm, const float* __restrict__ R, const float* __restrict__ t,
                                     float* __restrict__ PG) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;              // (row, point) with 3 x 96 points per row
  if (idx >= M * 288) return;
  const int row = idx / 288, pt = idx - row * 288;
  const float* p = Pm + (size_t)row * NPROJ + OFF_QP + pt * 3;
  const float* Rm = R + (size_t)row * 9;
  const float* tv = t + (size_t)row * 3;
  float* o = PG + (size_t)row * 864 + pt * 3;
  const float x = p[0], y = p[1], z = p[2];
  o[0] = Rm[0] * x + Rm[1] * y + Rm[2] * z + tv[0];
  o[1] = Rm[3] * x + Rm[4] * y + Rm[5] * z + tv[1];
  o[2] = Rm[6] * x + Rm[7] * y + Rm[8] * z + tv[2];
}
void bwd_points_global(int M, const float* Pm, const float* R, const float* t, float* PG, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  points_global_kernel<<<(M * 288 + 255) / 256, 256, 0, st>>>(M, Pm, R, t, PG);
}

// ------------------------------------------------------------------------------------------ aggregate backward, per row
// gfeat (M x 1824) -> g_agg (M x 288): gradient with respect to the aggregated GLOBAL points (before R^T (. - t)), from the
// point / norm / direction columns (ga.py:137-146; oracle/ipa_backward.py "aggregate backward").  The pair (cols 0..767) and
// node (768..1151) slices of gfeat are used in place by the pair kernels.
__global__ void aggr_bwd_kernel(int M, const float* __restrict__ gfeat, const float* __restrict__ feat, const float* __restrict__ R,
                                float* __restrict__ g_agg) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;              // (row, h * 8 + p)
  if (idx >= M * H * P) return;
  const int row = idx / (H * P), hp = idx - row * (H * P);
  const float* gf = gfeat + (size_t)row * NFEAT;
  const float* ff = feat + (size_t)row * NFEAT;
  float p[3], gp[3], gd[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) { p[a] = ff[FEAT_PTS + hp * 3 + a]; gp[a] = gf[FEAT_PTS + hp * 3 + a]; gd[a] = gf[FEAT_DIR + hp * 3 + a]; }
  const float nrm = ff[FEAT_DIST + hp], gn = gf[FEAT_DIST + hp];
  const float nc = fmaxf(nrm, 1e-30f), den = nrm + 1e-4f;
  const float dotgp = gd[0] * p[0] + gd[1] * p[1] + gd[2] * p[2];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float unit = p[a] / nc;
    gp[a] += gn * unit + gd[a] / den - unit * dotgp / (den * den);
  }
  const float* Rm = R + (size_t)row * 9;
  float* o = g_agg + (size_t)row * 288 + hp * 3;
#pragma unroll
  for (int a = 0; a < 3; ++a) o[a] = Rm[a * 3 + 0] * gp[0] + Rm[a * 3 + 1] * gp[1] + Rm[a * 3 + 2] * gp[2];      // pts = R^T (agg - t)
}
void bwd_aggregate(int M, const float* gfeat, const float* feat, const float* R, float* g_agg, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  aggr_bwd_kernel<<<(M * H * P + 255) / 256, 256, 0, st>>>(M, gfeat, feat, R, g_agg);
}

// ------------------------------------------------------------------------------------------ pairwise backward, query side
// One CTA per query row (b, i), two CTAs per SM.  Shared memory: the row block z[b,i,:,:] (L x 64), alpha and d logits of the row (L x 12 each)
// and the row's vectors.  Phases (threads over keys j unless noted):
//   1  d alpha[j][h] = g_p2n[h] . z[j] + g_node[h] . v[j,h] + g_agg[h] . vg[j,h]                     (ga.py:114-136 backward)
//      dot[h] = sum_j alpha d alpha
//   2  d logit[j][h] = alpha (d alpha - dot[h]) sqrt(1/3)  -> global (alpha layout), kept in smem    (softmax, ga.py:11-26,166)
//      partial d coef[h] = sum_j dl |qg_i - kg_j|^2 (per thread, block-reduced)
//   3  threads over outputs: d q[h][d] = sum_j dl k[j,h,d] / sqrt(32);  d qg[h][pc] = 2 c_h (qg_i sum_j dl - sum_j dl kg_j), rotated
//      to the local frame of residue i;  partial d W_b[h][c] = sum_j dl z[j][c]
//   4  threads over (j, c):  d z[j][c] (+)= sum_h alpha g_p2n[h][c] + sum_h dl W_b[h][c]
constexpr int PBQ_THREADS = 256;
constexpr int PBQ_HP = H + 1;            // row pitch of the per-key [12] arrays in shared memory (13: conflict-free over keys)
// (two [L][13] arrays, not three: with the per-key |qg - kg|^2 kept as well, the 116 KB of a 256-residue row were 576 bytes too many
// for two CTAs per SM, and this kernel lives on latency hiding: 8 -> 16 warps per SM)
static size_t pbq_smem(int L) { return ((size_t)L * C + 2 * (((size_t)L * PBQ_HP + 3) & ~size_t(3)) + 2 * H * C + H * D + 2 * H * P * 3 + 8 * H + 2 * H + 16) * sizeof(float); }

__global__ void __launch_bounds__(PBQ_THREADS) pair_bwd_query_kernel(const PairBwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int L = a.L, Lp = a.Lp;
  float* zs = sm;                              // [L][64]
  const size_t LH = ((size_t)L * PBQ_HP + 3) & ~size_t(3);      // keeps the float4-accessed arrays behind 16-byte aligned
  float* al = zs + (size_t)L * C;              // [L][13]
  float* dl = al + LH;                         // [L][13]  d alpha, then d logits
  float* gp2n = dl + LH;                       // [12][64]
  float* wb = gp2n + H * C;                    // [12][64]
  float* gnode = wb + H * C;                   // [12][32]
  float* gagg = gnode + H * D;                 // [12][24]  (phase 3: d query points, global frame)
  float* qgi = gagg + H * P * 3;               // [12][24]
  float* red = qgi + H * P * 3;                // [8 warps][12]
  float* dot = red + 8 * H;                    // [12]
  float* sdl = dot + H;                        // [12]
  const int row = blockIdx.x, b = row / L, i = row - b * L;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool live = a.mask[row] != 0;
  const float scale = 0.57735026918962576f;    // sqrt(1/3), ga.py:166
  // ---- stage the row
  const float4* zg = reinterpret_cast<const float4*>(a.z + (size_t)row * L * C);
  for (int e = tid; e < L * C / 4; e += PBQ_THREADS) reinterpret_cast<float4*>(zs)[e] = zg[e];
  for (int e = tid; e < L * H; e += PBQ_THREADS) {
    const int h = e / L, j = e - h * L;
    al[j * PBQ_HP + h] = live ? a.alpha[((size_t)(b * H + h) * L + i) * Lp + j] : 0.f;      // masked query rows: alpha = 0 (ga.py:25)
  }
  for (int e = tid; e < H * C; e += PBQ_THREADS) { gp2n[e] = a.gfeat[(size_t)row * NFEAT + e]; wb[e] = a.Wb[e]; }
  for (int e = tid; e < H * D; e += PBQ_THREADS) gnode[e] = a.gfeat[(size_t)row * NFEAT + FEAT_NODE + e];
  for (int e = tid; e < H * P * 3; e += PBQ_THREADS) { gagg[e] = a.g_agg[(size_t)row * 288 + e]; qgi[e] = a.PG[(size_t)row * 864 + e]; }
  __syncthreads();
  // ---- phase 1: d alpha and dot[h] = sum_j alpha d alpha
  float dotp[H];
#pragma unroll
  for (int h = 0; h < H; ++h) dotp[h] = 0.f;
  for (int j = tid; j < L; j += PBQ_THREADS) {
    const float* zr = zs + (size_t)j * C;
    const float* vj = a.Pm + ((size_t)b * L + j) * NPROJ + OFF_V;
    const float* vgj = a.PG + ((size_t)b * L + j) * 864 + 576;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float s = 0.f;
#pragma unroll 8
      for (int c = 0; c < C; ++c) { const int cc = (c + lane) & (C - 1); s = fmaf(gp2n[h * C + cc], zr[cc], s); }      // rotated: conflict-free
      for (int d = 0; d < D; d += 4) {
        const float4 v4 = *reinterpret_cast<const float4*>(vj + h * D + d);
        s = fmaf(gnode[h * D + d], v4.x, s); s = fmaf(gnode[h * D + d + 1], v4.y, s);
        s = fmaf(gnode[h * D + d + 2], v4.z, s); s = fmaf(gnode[h * D + d + 3], v4.w, s);
      }
      for (int c = 0; c < P * 3; c += 4) {
        const float4 v4 = *reinterpret_cast<const float4*>(vgj + h * P * 3 + c);
        s = fmaf(gagg[h * 24 + c], v4.x, s); s = fmaf(gagg[h * 24 + c + 1], v4.y, s);
        s = fmaf(gagg[h * 24 + c + 2], v4.z, s); s = fmaf(gagg[h * 24 + c + 3], v4.w, s);
      }
      dl[j * PBQ_HP + h] = s;
      dotp[h] = fmaf(al[j * PBQ_HP + h], s, dotp[h]);
    }
  }
#pragma unroll
  for (int h = 0; h < H; ++h) {
    const float v = wsum(dotp[h]);
    if (lane == 0) red[warp * H + h] = v;
  }
  __syncthreads();
  if (tid < H) {
    float s = 0.f;
    for (int w = 0; w < PBQ_THREADS / 32; ++w) s += red[w * H + tid];
    dot[tid] = s;
  }
  __syncthreads();
  // ---- phase 2: d logits (softmax backward); partial d coef[h] = sum_j d logit |qg_i - kg_j|^2 (per thread, reduced below)
  float dcp[H];
#pragma unroll
  for (int h = 0; h < H; ++h) dcp[h] = 0.f;
  for (int j = tid; j < Lp; j += PBQ_THREADS) {
    if (j >= L) {
#pragma unroll
      for (int h = 0; h < H; ++h) a.g_log[((size_t)(b * H + h) * L + i) * Lp + j] = 0.f;
      continue;
    }
    const float* kgj = a.PG + ((size_t)b * L + j) * 864 + 288;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float g = al[j * PBQ_HP + h] * (dl[j * PBQ_HP + h] - dot[h]) * scale;
      dl[j * PBQ_HP + h] = g;
      a.g_log[((size_t)(b * H + h) * L + i) * Lp + j] = g;
      float dd = 0.f;
      for (int c = 0; c < P * 3; c += 4) {
        const float4 k4 = *reinterpret_cast<const float4*>(kgj + h * 24 + c);
        const float e0 = qgi[h * 24 + c] - k4.x, e1 = qgi[h * 24 + c + 1] - k4.y, e2 = qgi[h * 24 + c + 2] - k4.z, e3 = qgi[h * 24 + c + 3] - k4.w;
        dd += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
      }
      dcp[h] = fmaf(g, dd, dcp[h]);
    }
  }
#pragma unroll
  for (int h = 0; h < H; ++h) {
    const float v = wsum(dcp[h]);
    if (lane == 0) red[warp * H + h] = v;                // (red was last read for dot[], two barriers ago)
  }
  __syncthreads();
  if (tid < H) {
    float s = 0.f;
    for (int j = 0; j < L; ++j) s += dl[j * PBQ_HP + tid];
    sdl[tid] = s;
    float c = 0.f;
    for (int w = 0; w < PBQ_THREADS / 32; ++w) c += red[w * H + tid];
    a.part[(size_t)row * 780 + 768 + tid] = c;
  }
  __syncthreads();
  // ---- phase 3: contractions over the keys, threads over outputs
  float* Grow = a.G + (size_t)row * NPROJ;
  for (int o = tid; o < H * D + H * P * 3 + H * C; o += PBQ_THREADS) {
    if (o < H * D) {                                   // d q[h][d] = sum_j dl k[j,h,d] / sqrt(32)
      const int h = o / D;
      const float* kcol = a.Pm + (size_t)b * L * NPROJ + OFF_K + o;
      float s = 0.f;
      for (int j = 0; j < L; ++j) s = fmaf(dl[j * PBQ_HP + h], kcol[(size_t)j * NPROJ], s);
      Grow[OFF_Q + o] = s * 0.17677669529663687f;
    } else if (o < H * D + H * P * 3) {                // d qg[h][pc] = 2 c_h (qg_i sum_j dl - sum_j dl kg_j)   (global frame)
      const int e = o - H * D, h = e / (P * 3);
      const float* kgcol = a.PG + (size_t)b * L * 864 + 288 + e;
      float s = 0.f;
      for (int j = 0; j < L; ++j) s = fmaf(dl[j * PBQ_HP + h], kgcol[(size_t)j * 864], s);
      gagg[e] = 2.f * a.coef[h] * (qgi[e] * sdl[h] - s);
    } else {                                           // partial d W_b[h][c] = sum_j dl z[j][c]
      const int e = o - H * D - H * P * 3, h = e / C, c = e - h * C;
      float s = 0.f;
      for (int j = 0; j < L; ++j) s = fmaf(dl[j * PBQ_HP + h], zs[(size_t)j * C + c], s);
      a.part[(size_t)row * 780 + e] = s;
    }
  }
  __syncthreads();
  {                                                    // query points: global -> local frame of residue i (q_global = R q_local + t)
    const float* Rm = a.R + (size_t)row * 9;
    for (int e = tid; e < H * P * 3; e += PBQ_THREADS) {
      const int pt = e / 3, c = e - pt * 3;
      Grow[OFF_QP + e] = Rm[0 * 3 + c] * gagg[pt * 3] + Rm[1 * 3 + c] * gagg[pt * 3 + 1] + Rm[2 * 3 + c] * gagg[pt * 3 + 2];
    }
  }
  // ---- phase 4: d z[j][c] (+)= sum_h alpha g_p2n[h][c] + sum_h dl W_b[h][c]
  float4* dzr = reinterpret_cast<float4*>(a.dz + (size_t)row * L * C);
  for (int e = tid; e < L * (C / 4); e += PBQ_THREADS) {
    const int j = e >> 4, c4 = (e & 15) * 4;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float av = al[j * PBQ_HP + h], gv = dl[j * PBQ_HP + h];
      const float4 p4 = *reinterpret_cast<const float4*>(gp2n + h * C + c4), w4 = *reinterpret_cast<const float4*>(wb + h * C + c4);
      o.x = fmaf(av, p4.x, fmaf(gv, w4.x, o.x)); o.y = fmaf(av, p4.y, fmaf(gv, w4.y, o.y));
      o.z = fmaf(av, p4.z, fmaf(gv, w4.z, o.z)); o.w = fmaf(av, p4.w, fmaf(gv, w4.w, o.w));
    }
    if (a.dz_accumulate) { const float4 old = dzr[e]; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
    dzr[e] = o;
  }
}

// ------------------------------------------------------------------------------------------ pairwise backward, key side
// The contractions over the QUERY index: one warp per (b, h, 32 keys), lane = key j; the block's four warps take four key tiles of
// the same (b, h) and share the staged per-query vectors (q, global query points, g_node, g_agg of head h: 112 floats per query).
//   d k[j]  = sum_i dl[i][j] q[i] / sqrt(32)            d kg[j] = -2 c_h (sum_i dl qg_i - kg_j sum_i dl)
//   d v[j]  = sum_i alpha[i][j] g_node[i]               d vg[j] = sum_i alpha[i][j] g_agg[i]
// (oracle/ipa_backward.py: g_k, g_kg, g_v, g_vg); point gradients are rotated to the local frame of residue j.
constexpr int PBK_WARPS = 4, PBK_I = 16;
__global__ void __launch_bounds__(PBK_WARPS * 32) pair_bwd_key_kernel(const PairBwdArgs a) {
  __shared__ float st[PBK_I][112];
  const int L = a.L, Lp = a.Lp;
  const int ntile = (L + 31) / 32, nblk = (ntile + PBK_WARPS - 1) / PBK_WARPS;
  const int bh = blockIdx.x / nblk, tb = blockIdx.x - bh * nblk;
  const int b = bh / H, h = bh - b * H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = (tb * PBK_WARPS + warp) * 32 + lane;
  const bool jok = j < L;
  float gk[D], gv[D], gkg[P * 3], gvg[P * 3], sgl = 0.f;
#pragma unroll
  for (int d = 0; d < D; ++d) { gk[d] = 0.f; gv[d] = 0.f; }
#pragma unroll
  for (int c = 0; c < P * 3; ++c) { gkg[c] = 0.f; gvg[c] = 0.f; }
  const float* arow = a.alpha + (size_t)(b * H + h) * L * Lp;
  const float* grow = a.g_log + (size_t)(b * H + h) * L * Lp;
  for (int i0 = 0; i0 < L; i0 += PBK_I) {
    __syncthreads();
    for (int e = threadIdx.x; e < PBK_I * 112; e += PBK_WARPS * 32) {
      const int ii = e / 112, c = e - ii * 112, i = i0 + ii;
      float v = 0.f;
      if (i < L) {
        const size_t r = (size_t)b * L + i;
        v = c < 32 ? a.Pm[r * NPROJ + OFF_Q + h * D + c]
                   : (c < 56 ? a.PG[r * 864 + h * 24 + (c - 32)]
                             : (c < 88 ? a.gfeat[r * NFEAT + FEAT_NODE + h * D + (c - 56)] : a.g_agg[r * 288 + h * 24 + (c - 88)]));
      }
      st[ii][c] = v;
    }
    __syncthreads();
    if (j < Lp) {
#pragma unroll 1
      for (int ii = 0; ii < PBK_I && i0 + ii < L; ++ii) {
        const float av = arow[(size_t)(i0 + ii) * Lp + j], gl = grow[(size_t)(i0 + ii) * Lp + j];
        sgl += gl;
#pragma unroll
        for (int d = 0; d < D; ++d) { gk[d] = fmaf(gl, st[ii][d], gk[d]); gv[d] = fmaf(av, st[ii][56 + d], gv[d]); }
#pragma unroll
        for (int c = 0; c < P * 3; ++c) { gkg[c] = fmaf(gl, st[ii][32 + c], gkg[c]); gvg[c] = fmaf(av, st[ii][88 + c], gvg[c]); }
      }
    }
  }
  if (!jok) return;
  const size_t row = (size_t)b * L + j;
  float* Grow = a.G + row * NPROJ;
#pragma unroll
  for (int d = 0; d < D; ++d) { Grow[OFF_K + h * D + d] = gk[d] * 0.17677669529663687f; Grow[OFF_V + h * D + d] = gv[d]; }
  const float* Rm = a.R + row * 9;
  const float* kgj = a.PG + row * 864 + 288 + h * 24;
  const float c2 = -2.f * a.coef[h];
#pragma unroll
  for (int p = 0; p < P; ++p) {
    float gg[3], vv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { gg[c] = c2 * (gkg[p * 3 + c] - kgj[p * 3 + c] * sgl); vv[c] = gvg[p * 3 + c]; }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      Grow[OFF_KP + h * 24 + p * 3 + c] = Rm[0 * 3 + c] * gg[0] + Rm[1 * 3 + c] * gg[1] + Rm[2 * 3 + c] * gg[2];
      Grow[OFF_VP + h * 24 + p * 3 + c] = Rm[0 * 3 + c] * vv[0] + Rm[1 * 3 + c] * vv[1] + Rm[2 * 3 + c] * vv[2];
    }
  }
}

// ------------------------------------------------------------------------------------------ EpsilonNet ends: inputs of the GEMMs
// cat0[row] = [res_feat | current_sequence_embedding[s_t]]   (dpm_full.py:86-88)
__global__ void mixer_cat_kernel(int M, const float* __restrict__ res_feat, const long long* __restrict__ s_t, const float* __restrict__ emb,
                                 float* __restrict__ cat0) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)M * 2 * F) return;
  const int r = (int)(i / (2 * F)), c = (int)(i - (size_t)r * 2 * F);
  long long s = s_t[r];
  s = s < 0 ? 0 : (s > 24 ? 24 : s);
  cat0[i] = c < F ? res_feat[(size_t)r * F + c] : emb[(size_t)s * F + (c - F)];
}
// hcat[row] = [x | beta, sin beta, cos beta]   (dpm_full.py:92-93)
__global__ void heads_cat_kernel(int M, int L, const float* __restrict__ x, const float* __restrict__ beta, float* __restrict__ hcat) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)M * (F + 3)) return;
  const int r = (int)(i / (F + 3)), c = (int)(i - (size_t)r * (F + 3));
  const float b = beta[r / L];
  hcat[i] = c < F ? x[(size_t)r * F + c] : (c == F ? b : (c == F + 1 ? sinf(b) : cosf(b)));
}
// d embedding[c] = sum over the rows with s_t = c of g_cat[row][128:256]; one CTA per class, fixed order
__global__ void embed_grad_kernel(int M, const long long* __restrict__ s_t, const float* __restrict__ g_cat, float* __restrict__ dE) {
  const int c = blockIdx.x, col = threadIdx.x;
  float s = 0.f;
  for (int r = 0; r < M; ++r) {
    long long k = s_t[r];
    k = k < 0 ? 0 : (k > 24 ? 24 : k);
    if ((int)k == c) s += g_cat[(size_t)r * 2 * F + F + col];
  }
  dE[(size_t)c * F + col] = s;
}
void bwd_mixer_cat(int M, const float* res_feat, const long long* s_t, const float* emb, float* cat0, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  mixer_cat_kernel<<<(unsigned)(((size_t)M * 2 * F + 255) / 256), 256, 0, st>>>(M, res_feat, s_t, emb, cat0);
}
void bwd_heads_cat(int M, int L, const float* x, const float* beta, float* hcat, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  heads_cat_kernel<<<(unsigned)(((size_t)M * (F + 3) + 255) / 256), 256, 0, st>>>(M, L, x, beta, hcat);
}
void bwd_embed_grad(int M, const long long* s_t, const float* g_cat, float* dE, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  embed_grad_kernel<<<25, F, 0, st>>>(M, s_t, g_cat, dE);
}

// ------------------------------------------------------------------------------------------ LayerNorm over 131 inputs (pRMSD head)
// forward: ln = (h - mu) / sqrt(var + 1e-10) * gamma + beta; backward as ln_bwd_kernel.  One warp per row, 131 = 4 x 32 + 3.
constexpr int F3 = F + 3;
__global__ void ln131_fwd_kernel(int M, const float* __restrict__ h, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* hr = h + (size_t)row * F3;
  float v[5], s = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) { const int c = lane + 32 * k; v[k] = c < F3 ? hr[c] : 0.f; s += v[k]; }
  const float mu = wsum(s) * (1.f / F3);
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) { const int c = lane + 32 * k; if (c < F3) q += (v[k] - mu) * (v[k] - mu); }
  const float inv = 1.f / sqrtf(wsum(q) * (1.f / F3) + 1e-10f);
#pragma unroll
  for (int k = 0; k < 5; ++k) { const int c = lane + 32 * k; if (c < F3) out[(size_t)row * F3 + c] = (v[k] - mu) * inv * gamma[c] + beta[c]; }
}
__global__ void ln131_bwd_kernel(int M, const float* __restrict__ g, const float* __restrict__ h, const float* __restrict__ gamma,
                                 float* __restrict__ dh_acc, float* __restrict__ gxh) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* hr = h + (size_t)row * F3;
  const float* gr = g + (size_t)row * F3;
  float v[5], gg[5], s = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) { const int c = lane + 32 * k; v[k] = c < F3 ? hr[c] : 0.f; gg[k] = c < F3 ? gr[c] : 0.f; s += v[k]; }
  const float mu = wsum(s) * (1.f / F3);
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) { const int c = lane + 32 * k; if (c < F3) q += (v[k] - mu) * (v[k] - mu); }
  const float inv = 1.f / sqrtf(wsum(q) * (1.f / F3) + 1e-10f);
  float xh[5], ax[5], s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int c = lane + 32 * k;
    xh[k] = c < F3 ? (v[k] - mu) * inv : 0.f;
    ax[k] = c < F3 ? gg[k] * gamma[c] : 0.f;
    s1 += ax[k]; s2 += ax[k] * xh[k];
  }
  const float m1 = wsum(s1) * (1.f / F3), m2 = wsum(s2) * (1.f / F3);
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int c = lane + 32 * k;
    if (c < F3) {
      dh_acc[(size_t)row * F3 + c] += inv * (ax[k] - m1 - xh[k] * m2);
      gxh[(size_t)row * F3 + c] = gg[k] * xh[k];
    }
  }
}
void bwd_ln131_fwd(int M, const float* h, const float* gamma, const float* beta, float* out, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  ln131_fwd_kernel<<<(M + 7) / 8, 256, 0, st>>>(M, h, gamma, beta, out);
}
void bwd_ln131_bwd(int M, const float* g, const float* h, const float* gamma, float* dh_acc, float* gxh, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  ln131_bwd_kernel<<<(M + 7) / 8, 256, 0, st>>>(M, g, h, gamma, dh_acc, gxh);
}

// ------------------------------------------------------------------------------------------ losses: derivatives wrt the head outputs
// (oracle/epsnet_backward.py "losses and their derivatives" / "heads backward"; dpm_full.py:186-232, :15-32, :369-378;
// transition.py:202-227).  stats kernel: counts and the pRMSD logit gradients; rows kernel: one thread per residue.
__global__ void loss_bwd_stats_kernel(LossBwdArgs a) {
  __shared__ double sh[256];
  const int M = a.N * a.L, tid = threadIdx.x;
  double ng = 0, dc = 0, m0 = 0;
  for (int r = tid; r < M; r += blockDim.x) { ng += a.mask_gen[r] ? 1.0 : 0.0; dc += (double)a.rows[(size_t)5 * M + r]; }
  for (int n = tid; n < a.N; n += blockDim.x) m0 += a.mask_gen[(size_t)n * a.L] ? 1.0 : 0.0;
  auto bsum = [&](double v) { __syncthreads(); sh[tid] = v; __syncthreads(); for (int s = blockDim.x / 2; s > 0; s >>= 1) { if (tid < s) sh[tid] += sh[tid + s]; __syncthreads(); } return sh[0]; };
  const double ngen = bsum(ng), dcnt = bsum(dc), m0s = bsum(m0);
  if (tid == 0) { a.stats[0] = (float)(ngen + 1e-8); a.stats[1] = (float)dcnt; a.stats[2] = (float)(m0s + 1e-10); }
  if (a.abdock && a.has_prmsd) {
    for (int n = tid; n < a.N; n += blockDim.x) {
      float s2 = 0.f, cnt = 0.f;
      for (int l = 0; l < a.L; ++l) { s2 += a.rows[(size_t)3 * M + (size_t)n * a.L + l]; cnt += a.mask_gen[(size_t)n * a.L + l] ? 1.f : 0.f; }
      const float rmsd = sqrtf(s2 / cnt);
      const float step = (a.dmax - a.dmin) / (float)(a.bins - 1);
      int best = 0; float bd = INFINITY;
      for (int k = 0; k < a.bins; ++k) {
        const float off = (k < a.bins / 2) ? a.dmin + step * (float)k : a.dmax - step * (float)(a.bins - 1 - k);     // torch.linspace
        const float d = fabsf(rmsd - off);
        if (d < bd) { bd = d; best = k; }
      }
      const float* lg = a.prmsd_logits + (size_t)n * a.bins;
      float mx = -INFINITY;
      for (int k = 0; k < a.bins; ++k) mx = fmaxf(mx, lg[k]);
      float se = 0.f;
      for (int k = 0; k < a.bins; ++k) se += expf(lg[k] - mx);
      const float mk = (a.mask_gen[(size_t)n * a.L] ? 1.f : 0.f) / (float)(m0s + 1e-10) * a.lw[3] / (float)a.L;      // ... / L: mean over ALL rows
      for (int k = 0; k < a.bins; ++k) a.glog[(size_t)n * a.bins + k] = (expf(lg[k] - mx) / se - (k == best ? 1.f : 0.f)) * mk;
    }
  }
}

// d (rotation of the normalised quaternion (1, b, c, d)) -> d (b, c, d)   (oracle/epsnet_backward.py:_quat_1ijk_backward)
__device__ __forceinline__ void quat1ijk_bwd(const float* o, const float* G, float* out) {
  const float s = sqrtf(1.f + o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
  const float a = 1.f / s, b = o[0] / s, c = o[1] / s, d = o[2] / s;
  const float tr = G[0] + G[4] + G[8];
  float gu[4];
  gu[0] = 2.f * (a * tr + d * (G[3] - G[1]) + c * (G[2] - G[6]) + b * (G[7] - G[5]));
  gu[1] = 2.f * (b * (G[0] - G[4] - G[8]) + c * (G[1] + G[3]) + d * (G[2] + G[6]) + a * (G[7] - G[5]));
  gu[2] = 2.f * (c * (G[4] - G[0] - G[8]) + b * (G[1] + G[3]) + a * (G[2] - G[6]) + d * (G[5] + G[7]));
  gu[3] = 2.f * (d * (G[8] - G[0] - G[4]) + a * (G[3] - G[1]) + b * (G[2] + G[6]) + c * (G[5] + G[7]));
  const float u[4] = {a, b, c, d};
  const float dt = u[0] * gu[0] + u[1] * gu[1] + u[2] * gu[2] + u[3] * gu[3];
#pragma unroll
  for (int k = 1; k < 4; ++k) out[k - 1] = (gu[k] - u[k] * dt) / s;
}

__global__ void __launch_bounds__(128) loss_bwd_rows_kernel(LossBwdArgs a, DiffW dw) {
  const int M = a.N * a.L;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  const int n = r / a.L;
  const int t = (int)a.tvec[n];
  const bool gen = a.mask_gen[r] != 0;
  float* go = a.GO + (size_t)r * 26;
#pragma unroll
  for (int k = 0; k < 26; ++k) go[k] = 0.f;
  if (a.GP) for (int k = 0; k < a.bins; ++k) a.GP[(size_t)r * a.bins + k] = a.glog[(size_t)n * a.bins + k];
  if (!gen) return;
  const float wgt = 1.f / a.stats[0];
  const float* Rm = a.R + (size_t)r * 9;
  // ---- rot: d (sum over columns of 1 - cos) / d R_pred, then R_pred = R U, U = quat(o_rot)
  {
    const Mat3 R0 = so3_exp(a.v_0[(size_t)r * 3], a.v_0[(size_t)r * 3 + 1], a.v_0[(size_t)r * 3 + 2]);
    const float* Rp = a.R_pred + (size_t)r * 9;
    float gR[9];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x0 = Rp[c], x1 = Rp[3 + c], x2 = Rp[6 + c];
      const float y0 = R0.m[c], y1 = R0.m[3 + c], y2 = R0.m[6 + c];
      const float nx2 = x0 * x0 + x1 * x1 + x2 * x2 + 1e-12f, ny2 = y0 * y0 + y1 * y1 + y2 * y2 + 1e-12f;
      const float rs = 1.f / sqrtf(nx2 * ny2);
      const float cs = (x0 * y0 + x1 * y1 + x2 * y2) * rs;
      const float w = wgt * a.lw[0];
      gR[c] = -(y0 * rs - x0 * cs / nx2) * w; gR[3 + c] = -(y1 * rs - x1 * cs / nx2) * w; gR[6 + c] = -(y2 * rs - x2 * cs / nx2) * w;
    }
    float gU[9];                                        // g_U[a][c] = sum_b R[b][a] g_Rpred[b][c]
#pragma unroll
    for (int x = 0; x < 3; ++x)
#pragma unroll
      for (int c = 0; c < 3; ++c) gU[x * 3 + c] = Rm[0 * 3 + x] * gR[c] + Rm[1 * 3 + x] * gR[3 + c] + Rm[2 * 3 + x] * gR[6 + c];
    quat1ijk_bwd(a.o_rot + (size_t)r * a.ld_orot, gU, go + 3);
  }
  // ---- pos (+ dist): d / d eps_pos, then eps_pos = R o_crd on generated residues
  {
    float p0[3], pn[3], pp[3], ge[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      p0[i] = (a.p_0_ang[(size_t)r * 3 + i] - dw.pos_mean[i]) / dw.pos_scale;
      pn[i] = (a.p_noisy_ang[(size_t)r * 3 + i] - dw.pos_mean[i]) / dw.pos_scale;
      pp[i] = a.eps_pos[(size_t)r * 3 + i];
      const float tgt = a.abdock ? (a.pred_x0 ? p0[i] : pn[i]) : (a.z ? a.z[(size_t)r * 3 + i] : 0.f);
      ge[i] = 2.f * (pp[i] - tgt) * wgt * a.lw[1];
    }
    if (a.abdock && a.pred_x0 && a.mask_res[r] && a.stats[1] > 0.f) {
      const int base_r = n * a.L;
      const float inv_cnt = a.lw[4] / a.stats[1];
      for (int j = 0; j < a.L; ++j) {
        if (!a.mask_res[base_r + j]) continue;
        float dv[3], dp = 0.f, dt = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          dv[i] = pp[i] - a.eps_pos[(size_t)(base_r + j) * 3 + i];
          const float q0 = (a.p_0_ang[(size_t)(base_r + j) * 3 + i] - dw.pos_mean[i]) / dw.pos_scale;
          dp += dv[i] * dv[i]; dt += (p0[i] - q0) * (p0[i] - q0);
        }
        dp = sqrtf(dp); dt = sqrtf(dt);
        if (!(dp > 0.f)) continue;
        const float u = dp - dt;
        const float hp = fabsf(u) < 1.f ? u : (u > 0.f ? 1.f : -1.f);
        const float nsel = 1.f + (a.mask_gen[base_r + j] ? 1.f : 0.f);      // the pair counts once per generated end (row i, and row j if generated)
        const float cf = hp * nsel * inv_cnt / dp;
#pragma unroll
        for (int i = 0; i < 3; ++i) ge[i] += cf * dv[i];
      }
    }
#pragma unroll
    for (int x = 0; x < 3; ++x) go[x] = Rm[0 * 3 + x] * ge[0] + Rm[1 * 3 + x] * ge[1] + Rm[2 * 3 + x] * ge[2];      // g_crd = R^T g_eps_pos
  }
  // ---- seq: KL(posterior(s_t, s_0) || posterior(s_t, softmax(o_seq)))
  {
    const float ab = dw.alpha_bars_seq[t];
    const float base = (1.f - ab) / (float)NAA;
    const long long st = a.s_noisy[r], s0 = a.s_0[r];
    float At[NAA], th[NAA], pt[NAA], ssum = 0.f, tsum = 0.f;
#pragma unroll
    for (int k = 0; k < NAA; ++k) {
      const float ct = (st == k) ? 1.f : 0.f, c0 = (s0 == k) ? 1.f : 0.f;
      At[k] = ab * ct + base;
      pt[k] = At[k] * (ab * c0 + base);
      th[k] = At[k] * (ab * a.c_den[(size_t)r * NAA + k] + base);
      ssum += th[k]; tsum += pt[k];
    }
    const float S = ssum + 1e-8f;
    float gth[NAA], gdot = 0.f;
#pragma unroll
    for (int k = 0; k < NAA; ++k) {
      const float post_true = pt[k] / (tsum + 1e-8f), post_pred = th[k] / S;
      gth[k] = -post_true / (post_pred + 1e-8f) * wgt * a.lw[2];      // d / d post_pred
      gdot += gth[k] * th[k];
    }
    float gc[NAA], cd = 0.f;
#pragma unroll
    for (int k = 0; k < NAA; ++k) {
      const float gt = gth[k] / S - gdot / (S * S);                    // d / d theta
      gc[k] = gt * At[k] * ab;                                         // d / d c_denoised
      cd += a.c_den[(size_t)r * NAA + k] * gc[k];
    }
#pragma unroll
    for (int k = 0; k < NAA; ++k) go[6 + k] = a.c_den[(size_t)r * NAA + k] * (gc[k] - cd);      // softmax backward
  }
}
void launch_loss_bwd(const LossBwdArgs& a, const DiffW& dw, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  loss_bwd_stats_kernel<<<1, 256, 0, st>>>(a);
  count_launch();
  loss_bwd_rows_kernel<<<(a.N * a.L + 127) / 128, 128, 0, st>>>(a, dw);
}

cudaError_t backward_kernels_init() {
  return cudaFuncSetAttribute(pair_bwd_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pbq_smem(ABOPT_MAX_L_INTERNAL));
}
void launch_pair_bwd(const PairBwdArgs& a, cudaStream_t st) {
  {
    ProfScope prof__(KK_PAIR, st);
    pair_bwd_query_kernel<<<a.N * a.L, PBQ_THREADS, pbq_smem(a.L), st>>>(a);
  }
  {
    ProfScope prof__(KK_AGGR, st);
    const int ntile = (a.L + 31) / 32, nblk = (ntile + PBK_WARPS - 1) / PBK_WARPS;
    pair_bwd_key_kernel<<<a.N * H * nblk, PBK_WARPS * 32, 0, st>>>(a);
  }
}

}  // namespace abopt
