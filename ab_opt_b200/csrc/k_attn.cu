// Pairwise part of one GABlock, decoupled into three kernels that exchange the (heads x L x L)
// logits / attention weights through L2 (the batch is processed in chunks small enough to stay
// L2-resident, see api.cu):
//   logits_kernel      : node + spatial logits, a register-tiled batched "Q K^T", + pair bias, scale,
//                        key mask -> final logits                                 ga.py:81-86,92-112,166,23
//   pair_stream_kernel : (k_pair.cu) streams z ONCE per layer: softmax over j, pair aggregation
//   aggr_kernel        : node + point aggregation ("P V") and the local-frame features   ga.py:120-147
// Layouts: S / alpha are [chunk complex][head][i][Lp] (j contiguous, Lp = L rounded up to 4).
#include "common.cuh"
#include "params.cuh"
#include "kernels.h"

namespace abopt {

// ------------------------------------------------------------------------------------------ logits
constexpr int LG_T = 64;          // tile edge (i and j)
constexpr int LG_K = D + P * 3;   // 56 = 32 qk channels + 24 point coordinates
constexpr int LG_LD = LG_T + 4;

__global__ void __launch_bounds__(256, 2)
logits_kernel(int L, int Lp, const float* __restrict__ proj, const float* __restrict__ coef, const float* __restrict__ bias,
              const uint8_t* __restrict__ mask, float* __restrict__ S) {
  __shared__ __align__(16) float Qs[LG_K][LG_LD];
  __shared__ __align__(16) float Ks[LG_K][LG_LD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int bh = blockIdx.z, b = bh / H, h = bh % H;
  const int i0 = blockIdx.y * LG_T, j0 = blockIdx.x * LG_T;

  // 64 rows x 14 float4 per operand; transposed into [d][row]
  for (int f = tid; f < LG_T * (LG_K / 4); f += 256) {
    const int r = f / (LG_K / 4), q4 = f % (LG_K / 4);
    const int d = q4 * 4;
    const int qoff = (d < D) ? (OFF_Q + h * D + d) : (OFF_QP + h * P * 3 + (d - D));
    const int koff = (d < D) ? (OFF_K + h * D + d) : (OFF_KP + h * P * 3 + (d - D));
    float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), kv = qv;
    if (i0 + r < L) qv = *reinterpret_cast<const float4*>(proj + (size_t)(b * L + i0 + r) * NPROJ + qoff);
    if (j0 + r < L) kv = *reinterpret_cast<const float4*>(proj + (size_t)(b * L + j0 + r) * NPROJ + koff);
    Qs[d + 0][r] = qv.x; Qs[d + 1][r] = qv.y; Qs[d + 2][r] = qv.z; Qs[d + 3][r] = qv.w;
    Ks[d + 0][r] = kv.x; Ks[d + 1][r] = kv.y; Ks[d + 2][r] = kv.z; Ks[d + 3][r] = kv.w;
  }
  __syncthreads();

  float nd[4][4], sp[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) { nd[a][c] = 0.f; sp[a][c] = 0.f; }
#pragma unroll 8
  for (int d = 0; d < D; ++d) {
    const float4 q = *reinterpret_cast<const float4*>(&Qs[d][ty * 4]);
    const float4 k = *reinterpret_cast<const float4*>(&Ks[d][tx * 4]);
    const float qa[4] = {q.x, q.y, q.z, q.w}, ka[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) nd[a][c] = fmaf(qa[a], ka[c], nd[a][c]);
  }
#pragma unroll 8
  for (int d = D; d < LG_K; ++d) {
    const float4 q = *reinterpret_cast<const float4*>(&Qs[d][ty * 4]);
    const float4 k = *reinterpret_cast<const float4*>(&Ks[d][tx * 4]);
    const float qa[4] = {q.x, q.y, q.z, q.w}, ka[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) { const float df = qa[a] - ka[c]; sp[a][c] = fmaf(df, df, sp[a][c]); }
  }
  const float cf = coef[h];
  const float inv_sqrt_d = 0.17677669529663687f;      // 1/sqrt(32)
  const float scale = 0.57735026918962576f;           // sqrt(1/3), ga.py:166
  // + pair bias (pair_bias_kernel), * sqrt(1/3), key mask as a finite -1e5 (ga.py:23,166): the final logits
  const int jq = j0 + tx * 4;
  float pen[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) pen[c] = (jq + c < L && mask[(size_t)b * L + jq + c] != 0) ? 0.f : 1e5f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ty * 4 + a;
    if (i < L && jq < L) {
      const size_t off = ((size_t)bh * L + i) * Lp + jq;
      const float4 pbv = __ldg(reinterpret_cast<const float4*>(bias + off));
      const float pbs[4] = {pbv.x, pbv.y, pbv.z, pbv.w};
      float o[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) o[c] = ((nd[a][c] * inv_sqrt_d + pbs[c]) + sp[a][c] * cf) * scale - pen[c];
      *reinterpret_cast<float4*>(S + off) = make_float4(o[0], o[1], o[2], o[3]);      // columns >= L are padding
    }
  }
}

// ------------------------------------------------------------------------------------------ aggregation
constexpr int AG_TI = 64, AG_TJ = 32, AG_N = D + P * 3;      // 56 value columns per head
constexpr int AG_ALD = AG_TI + 4, AG_VLD = AG_N, AG_OLD = AG_N + 1;

__global__ void __launch_bounds__(256, 2)
aggr_kernel(int L, int Lp, int b0, const float* __restrict__ alpha, const float* __restrict__ proj,
            const float* __restrict__ R, const float* __restrict__ t, float* __restrict__ feat, float* __restrict__ feat_lo) {
  __shared__ __align__(16) float As[AG_TJ][AG_ALD];      // alpha tile, transposed [j][i]
  __shared__ __align__(16) float Vs[AG_TJ][AG_VLD];      // [j][n]  n < 32: value channels, n >= 32: global value points
  __shared__ float Os[AG_TI][AG_OLD];
  const int tid = threadIdx.x;
  const int bh = blockIdx.y, bl = bh / H, h = bh % H, b = b0 + bl;
  const int i0 = blockIdx.x * AG_TI;
  const int tx = tid % 14, ty = tid / 14;                // 14 column groups x 16 row groups (224 threads)
  const bool active = tid < 224;

  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;

  for (int j0 = 0; j0 < L; j0 += AG_TJ) {
    // alpha[h][i0..i0+63][j0..j0+31]: 64 x 8 float4
    for (int f = tid; f < AG_TI * (AG_TJ / 4); f += 256) {
      const int r = f >> 3, q4 = f & 7;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i0 + r < L && j0 + q4 * 4 < Lp) v = *reinterpret_cast<const float4*>(alpha + ((size_t)(bl * H + h) * L + i0 + r) * Lp + j0 + q4 * 4);
      As[q4 * 4 + 0][r] = v.x; As[q4 * 4 + 1][r] = v.y; As[q4 * 4 + 2][r] = v.z; As[q4 * 4 + 3][r] = v.w;
    }
    // values: 32 rows x 14 float4
    for (int f = tid; f < AG_TJ * (AG_N / 4); f += 256) {
      const int r = f / (AG_N / 4), q4 = f % (AG_N / 4);
      const int n = q4 * 4;
      const int coff = (n < D) ? (OFF_V + h * D + n) : (OFF_VP + h * P * 3 + (n - D));
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + r < L) v = *reinterpret_cast<const float4*>(proj + (size_t)(b * L + j0 + r) * NPROJ + coff);
      *reinterpret_cast<float4*>(&Vs[r][n]) = v;
    }
    __syncthreads();
    if (active) {
#pragma unroll 8
      for (int j = 0; j < AG_TJ; ++j) {
        const float4 a = *reinterpret_cast<const float4*>(&As[j][ty * 4]);
        const float4 v = *reinterpret_cast<const float4*>(&Vs[j][tx * 4]);
        const float aa[4] = {a.x, a.y, a.z, a.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[p][c] = fmaf(aa[p], vv[c], acc[p][c]);
      }
    }
    __syncthreads();
  }
  if (active)
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int c = 0; c < 4; ++c) Os[ty * 4 + p][tx * 4 + c] = acc[p][c];
  __syncthreads();

  // node aggregate -> feat[:, 768 + h*32 + d]   (ga.py:120-125)
  for (int o = tid; o < AG_TI * D; o += 256) {
    const int r = o >> 5, d = o & 31;
    if (i0 + r < L) {
      const size_t o2 = ((size_t)b * L + i0 + r) * NFEAT + FEAT_NODE + h * D + d;
      feat[o2] = Os[r][d]; feat_lo[o2] = tf32_lo(Os[r][d]);
    }
  }
  // point aggregate -> local frame, norm, direction   (ga.py:137-146)
  for (int o = tid; o < AG_TI * P; o += 256) {
    const int r = o >> 3, p = o & 7;
    const int i = i0 + r;
    if (i >= L) continue;
    const size_t row = (size_t)b * L + i;
    float Rm[9], tv[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) Rm[k] = __ldg(R + row * 9 + k);
#pragma unroll
    for (int k = 0; k < 3; ++k) tv[k] = __ldg(t + row * 3 + k);
    const float gx = Os[r][D + p * 3 + 0] - tv[0], gy = Os[r][D + p * 3 + 1] - tv[1], gz = Os[r][D + p * 3 + 2] - tv[2];
    // p = R^T (q - t)   (geometry.py:94-113)
    const float lx = Rm[0] * gx + Rm[3] * gy + Rm[6] * gz;
    const float ly = Rm[1] * gx + Rm[4] * gy + Rm[7] * gz;
    const float lz = Rm[2] * gx + Rm[5] * gy + Rm[8] * gz;
    const float nrm = sqrtf(lx * lx + ly * ly + lz * lz);
    const float den = nrm + 1e-4f;                        // normalize_vector(eps=1e-4), ga.py:139
    float* fr = feat + row * NFEAT;
    float* fl = feat_lo + row * NFEAT;
    const int hp = h * P + p;
    const float dx = lx / den, dy = ly / den, dz = lz / den;
    fr[FEAT_PTS + hp * 3 + 0] = lx; fr[FEAT_PTS + hp * 3 + 1] = ly; fr[FEAT_PTS + hp * 3 + 2] = lz;
    fr[FEAT_DIST + hp] = nrm;
    fr[FEAT_DIR + hp * 3 + 0] = dx; fr[FEAT_DIR + hp * 3 + 1] = dy; fr[FEAT_DIR + hp * 3 + 2] = dz;
    fl[FEAT_PTS + hp * 3 + 0] = tf32_lo(lx); fl[FEAT_PTS + hp * 3 + 1] = tf32_lo(ly); fl[FEAT_PTS + hp * 3 + 2] = tf32_lo(lz);
    fl[FEAT_DIST + hp] = tf32_lo(nrm);
    fl[FEAT_DIR + hp * 3 + 0] = tf32_lo(dx); fl[FEAT_DIR + hp * 3 + 1] = tf32_lo(dy); fl[FEAT_DIR + hp * 3 + 2] = tf32_lo(dz);
  }
}

// ------------------------------------------------------------------------------------------ taps
// alpha [chunk][h][i][Lp]  ->  reference layout (N, L, L, 12)   (parity taps only)
__global__ void alpha_to_reference_layout(int L, int Lp, int b0, const float* __restrict__ alpha, float* __restrict__ out) {
  const size_t n = (size_t)gridDim.y * L * L * H;
  (void)n;
  const int bl = blockIdx.y;
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < (size_t)L * L * H; o += (size_t)gridDim.x * blockDim.x) {
    const int h = o % H;
    const size_t ij = o / H;
    const int j = ij % L, i = ij / L;
    out[(size_t)(b0 + bl) * L * L * H + o] = alpha[((size_t)(bl * H + h) * L + i) * Lp + j];
  }
}

// ------------------------------------------------------------------------------------------ launchers
cudaError_t attn_kernels_init() { return cudaSuccess; }

void launch_logits(int nb, int L, int Lp, const float* proj_chunk, const float* coef, const float* bias_chunk,
                   const uint8_t* mask_chunk, float* S, cudaStream_t st) {
  ProfScope prof__(KK_LOGITS, st);
  dim3 grid((L + LG_T - 1) / LG_T, (L + LG_T - 1) / LG_T, nb * H);
  logits_kernel<<<grid, 256, 0, st>>>(L, Lp, proj_chunk, coef, bias_chunk, mask_chunk, S);
}

void launch_aggr(int nb, int b0, int L, int Lp, const float* alpha, const float* proj, const float* R, const float* t,
                 float* feat, float* feat_lo, cudaStream_t st) {
  ProfScope prof__(KK_AGGR, st);
  dim3 grid((L + AG_TI - 1) / AG_TI, nb * H);
  aggr_kernel<<<grid, 256, 0, st>>>(L, Lp, b0, alpha, proj, R, t, feat, feat_lo);
}

void launch_alpha_tap(int nb, int b0, int L, int Lp, const float* alpha, float* out, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  dim3 grid(64, nb);
  alpha_to_reference_layout<<<grid, 256, 0, st>>>(L, Lp, b0, alpha, out);
}

}  // namespace abopt
