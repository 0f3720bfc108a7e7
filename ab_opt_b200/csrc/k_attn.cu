// Pairwise part of one GABlock, decoupled into three kernels that exchange the (heads x L x L)
// logits / attention weights through L2 (the batch is processed in chunks small enough to stay
// L2-resident, see api.cu):
//   logits_kernel : node + spatial logits, a register-tiled batched "Q K^T"         ga.py:81-86,92-112
//   pair_kernel   : streams z ONCE per layer: pair bias, masked softmax over j,
//                   pair aggregation                                               ga.py:88-90,11-26,114-118
//   aggr_kernel   : node + point aggregation ("P V") and the local-frame features   ga.py:120-147
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
logits_kernel(int L, int Lp, const float* __restrict__ proj, const float* __restrict__ coef, float* __restrict__ S) {
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
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ty * 4 + a, j = j0 + tx * 4;
    if (i < L && j < L) {
      float4 o = make_float4(nd[a][0] * inv_sqrt_d + sp[a][0] * cf, nd[a][1] * inv_sqrt_d + sp[a][1] * cf,
                             nd[a][2] * inv_sqrt_d + sp[a][2] * cf, nd[a][3] * inv_sqrt_d + sp[a][3] * cf);
      *reinterpret_cast<float4*>(S + ((size_t)bh * L + i) * Lp + j) = o;      // columns >= L are padding
    }
  }
}

// ------------------------------------------------------------------------------------------ pair stream
// One CTA per query residue (b, i).  The residue's z row-block (L x 64 floats) is staged in shared
// memory once, XOR-swizzled at 16-byte granularity so that both access patterns are conflict free:
//   (a) thread = key residue j reads its own 256 B row          (pair bias, z . Wb)
//   (b) 16 lanes = the 16 float4 column groups of one row j      (pair aggregation, alpha . z)
constexpr int PK_THREADS = 256;
constexpr int PK_SLICES = PK_THREADS / 16;       // 16 j-slices in the aggregation phase

template <int JPT>   // key residues per thread: L <= 256 * JPT
__global__ void __launch_bounds__(PK_THREADS, 2)
pair_kernel(int L, int Lp, int b0, const float* __restrict__ z, const uint8_t* __restrict__ mask,
            const float* __restrict__ S, const __grid_constant__ PairBiasParams pb,
            float* __restrict__ alpha, float* __restrict__ feat) {
  extern __shared__ __align__(16) float smem[];
  const int zs_floats = max(L * C, PK_SLICES * H * C);
  float* zs = smem;                       // [L][64] swizzled; later the cross-slice reduction buffer
  float* al = smem + zs_floats;           // [L][12]
  __shared__ float red[PK_THREADS / 32][H];
  __shared__ float fin[H];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bl = blockIdx.x / L, i = blockIdx.x % L;       // complex within the chunk, query residue
  const int b = b0 + bl;
  const bool row_ok = mask[(size_t)b * L + i] != 0;
  float* feat_row = feat + ((size_t)b * L + i) * NFEAT;

  if (!row_ok) {
    // masked query: alpha row = 0 (ga.py:25) -> zero pair aggregate
    for (int o = tid; o < H * C; o += PK_THREADS) feat_row[o] = 0.f;
    for (int o = tid; o < H * Lp; o += PK_THREADS) {
      const int h = o / Lp, j = o % Lp;
      alpha[((size_t)(bl * H + h) * L + i) * Lp + j] = 0.f;
    }
    return;
  }

  // ---- stage z[b, i, :, :] (coalesced 16 B cp.async, swizzled destination)
  const float* zrow = z + ((size_t)b * L + i) * (size_t)L * C;
  for (int ch = tid; ch < L * 16; ch += PK_THREADS) {
    const int j = ch >> 4, q = ch & 15;
    cp_async16(zs + j * C + ((q ^ (j & 15)) << 2), zrow + (size_t)ch * 4);
  }
  cp_async_commit();

  // ---- logits of this thread's key residues: S (node + spatial, from L2) while z is in flight
  float lg[JPT][H];
  bool jok[JPT];
#pragma unroll
  for (int u = 0; u < JPT; ++u) {
    const int j = tid + u * PK_THREADS;
    jok[u] = (j < L);
#pragma unroll
    for (int h = 0; h < H; ++h)
      lg[u][h] = jok[u] ? __ldg(S + ((size_t)(bl * H + h) * L + i) * Lp + j) : 0.f;
  }
  cp_async_wait<0>();
  __syncthreads();

  const float scale = 0.57735026918962576f;      // sqrt(1/3), ga.py:166
#pragma unroll
  for (int u = 0; u < JPT; ++u) {
    const int j = tid + u * PK_THREADS;
    if (jok[u]) {
      float bias[H];
#pragma unroll
      for (int h = 0; h < H; ++h) bias[h] = 0.f;
      const float* zr = zs + j * C;
      const int sw = j & 15;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(zr + ((q ^ sw) << 2));
#pragma unroll
        for (int h = 0; h < H; ++h) {
          bias[h] = fmaf(v.x, pb.Wb[q * 4 + 0][h], bias[h]);
          bias[h] = fmaf(v.y, pb.Wb[q * 4 + 1][h], bias[h]);
          bias[h] = fmaf(v.z, pb.Wb[q * 4 + 2][h], bias[h]);
          bias[h] = fmaf(v.w, pb.Wb[q * 4 + 3][h], bias[h]);
        }
      }
      const bool mj = mask[(size_t)b * L + j] != 0;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float v = (lg[u][h] + bias[h]) * scale;
        lg[u][h] = mj ? v : v - 1e5f;                 // ga.py:23 (finite "-inf")
      }
    } else {
#pragma unroll
      for (int h = 0; h < H; ++h) lg[u][h] = -INFINITY;
    }
  }

  // ---- softmax over j (ga.py:24), block-wide per head
  float mx[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    float m = lg[0][h];
#pragma unroll
    for (int u = 1; u < JPT; ++u) m = fmaxf(m, lg[u][h]);
    mx[h] = warp_max(m);
  }
  if (lane == 0)
#pragma unroll
    for (int h = 0; h < H; ++h) red[warp][h] = mx[h];
  __syncthreads();
  if (tid < H) {
    float m = red[0][tid];
    for (int w = 1; w < PK_THREADS / 32; ++w) m = fmaxf(m, red[w][tid]);
    fin[tid] = m;
  }
  __syncthreads();
  float sm[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    const float m = fin[h];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < JPT; ++u) { lg[u][h] = expf(lg[u][h] - m); s += lg[u][h]; }   // exp(-inf) = 0 for j >= L
    sm[h] = warp_sum(s);
  }
  __syncthreads();                                      // everyone has read fin[] (max)
  if (lane == 0)
#pragma unroll
    for (int h = 0; h < H; ++h) red[warp][h] = sm[h];
  __syncthreads();
  if (tid < H) {
    float s = 0.f;
    for (int w = 0; w < PK_THREADS / 32; ++w) s += red[w][tid];
    fin[tid] = s;
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < JPT; ++u) {
    const int j = tid + u * PK_THREADS;
    if (j < Lp) {
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float a = (j < L) ? lg[u][h] / fin[h] : 0.f;
        lg[u][h] = a;
        alpha[((size_t)(bl * H + h) * L + i) * Lp + j] = a;
      }
      if (j < L) {
        float4* dst = reinterpret_cast<float4*>(al + j * H);
        dst[0] = make_float4(lg[u][0], lg[u][1], lg[u][2], lg[u][3]);
        dst[1] = make_float4(lg[u][4], lg[u][5], lg[u][6], lg[u][7]);
        dst[2] = make_float4(lg[u][8], lg[u][9], lg[u][10], lg[u][11]);
      }
    }
  }
  __syncthreads();

  // ---- pair aggregation out[h][c] = sum_j alpha[j][h] z[j][c]   (ga.py:114-118)
  const int c4 = tid & 15, js = tid >> 4;
  float acc[H][4];
#pragma unroll
  for (int h = 0; h < H; ++h) { acc[h][0] = acc[h][1] = acc[h][2] = acc[h][3] = 0.f; }
  for (int j = js; j < L; j += PK_SLICES) {
    const float4 zv = *reinterpret_cast<const float4*>(zs + j * C + ((c4 ^ (j & 15)) << 2));
    const float4* ap = reinterpret_cast<const float4*>(al + j * H);
    const float4 a0 = ap[0], a1 = ap[1], a2 = ap[2];
    const float a[H] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
#pragma unroll
    for (int h = 0; h < H; ++h) {
      acc[h][0] = fmaf(a[h], zv.x, acc[h][0]); acc[h][1] = fmaf(a[h], zv.y, acc[h][1]);
      acc[h][2] = fmaf(a[h], zv.z, acc[h][2]); acc[h][3] = fmaf(a[h], zv.w, acc[h][3]);
    }
  }
  __syncthreads();                                      // all reads of zs done -> reuse as reduction buffer
#pragma unroll
  for (int h = 0; h < H; ++h)
    *reinterpret_cast<float4*>(zs + js * (H * C) + h * C + c4 * 4) = make_float4(acc[h][0], acc[h][1], acc[h][2], acc[h][3]);
  __syncthreads();
  for (int o = tid; o < H * C; o += PK_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < PK_SLICES; ++k) s += zs[k * (H * C) + o];
    feat_row[o] = s;
  }
}

// ------------------------------------------------------------------------------------------ aggregation
constexpr int AG_TI = 64, AG_TJ = 32, AG_N = D + P * 3;      // 56 value columns per head
constexpr int AG_ALD = AG_TI + 4, AG_VLD = AG_N, AG_OLD = AG_N + 1;

__global__ void __launch_bounds__(256, 2)
aggr_kernel(int L, int Lp, int b0, const float* __restrict__ alpha, const float* __restrict__ proj,
            const float* __restrict__ R, const float* __restrict__ t, float* __restrict__ feat) {
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
    if (i0 + r < L) feat[((size_t)b * L + i0 + r) * NFEAT + FEAT_NODE + h * D + d] = Os[r][d];
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
    const int hp = h * P + p;
    fr[FEAT_PTS + hp * 3 + 0] = lx; fr[FEAT_PTS + hp * 3 + 1] = ly; fr[FEAT_PTS + hp * 3 + 2] = lz;
    fr[FEAT_DIST + hp] = nrm;
    fr[FEAT_DIR + hp * 3 + 0] = lx / den; fr[FEAT_DIR + hp * 3 + 1] = ly / den; fr[FEAT_DIR + hp * 3 + 2] = lz / den;
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
size_t pair_smem_bytes(int L) {
  const size_t zs = (size_t)((L * C > PK_SLICES * H * C) ? L * C : PK_SLICES * H * C);
  return (zs + (size_t)L * H) * sizeof(float);
}

cudaError_t attn_kernels_init() {
  cudaError_t e;
  const int mx = 200 * 1024;
  if ((e = cudaFuncSetAttribute(pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(pair_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(pair_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024)) != cudaSuccess) return e;
  return cudaSuccess;
}

void launch_logits(int nb, int L, int Lp, const float* proj_chunk, const float* coef, float* S, cudaStream_t st) {
  ProfScope prof__(KK_LOGITS, st);
  dim3 grid((L + LG_T - 1) / LG_T, (L + LG_T - 1) / LG_T, nb * H);
  logits_kernel<<<grid, 256, 0, st>>>(L, Lp, proj_chunk, coef, S);
}

void launch_pair(int nb, int b0, int L, int Lp, const float* z, const uint8_t* mask, const float* S,
                 const PairBiasParams& pb, float* alpha, float* feat, cudaStream_t st) {
  ProfScope prof__(KK_PAIR, st);
  const size_t smem = pair_smem_bytes(L);
  const int grid = nb * L;
  if (L <= 256) pair_kernel<1><<<grid, PK_THREADS, smem, st>>>(L, Lp, b0, z, mask, S, pb, alpha, feat);
  else if (L <= 512) pair_kernel<2><<<grid, PK_THREADS, smem, st>>>(L, Lp, b0, z, mask, S, pb, alpha, feat);
  else pair_kernel<3><<<grid, PK_THREADS, smem, st>>>(L, Lp, b0, z, mask, S, pb, alpha, feat);
}

void launch_aggr(int nb, int b0, int L, int Lp, const float* alpha, const float* proj, const float* R, const float* t,
                 float* feat, cudaStream_t st) {
  ProfScope prof__(KK_AGGR, st);
  dim3 grid((L + AG_TI - 1) / AG_TI, nb * H);
  aggr_kernel<<<grid, 256, 0, st>>>(L, Lp, b0, alpha, proj, R, t, feat);
}

void launch_alpha_tap(int nb, int b0, int L, int Lp, const float* alpha, float* out, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  dim3 grid(64, nb);
  alpha_to_reference_layout<<<grid, 256, 0, st>>>(L, Lp, b0, alpha, out);
}

}  // namespace abopt
