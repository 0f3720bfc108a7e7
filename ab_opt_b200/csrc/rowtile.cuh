// Row-tile GEMM machinery: a CTA of 256 threads owns 8*R rows x 128 output columns (R = rows per warp: 8 for the
// whole-batch launches, fewer for the short row lists of the sampling loop, where the number of CTAs is what matters).
//   thread (warp w, lane l) accumulates rows  w*R .. w*R+R-1  and columns  l*4 .. l*4+3,
//   so a warp holds R complete rows and row-wise reductions (LayerNorm) are warp shuffles.
// R is deduced from the accumulator array.  Every output element is the same chain of FMAs in the same k order whatever R is,
// so the results do not depend on it.
// Operand A is row-major in shared memory (rows x k, k contiguous); operand B is a weight stored
// K-major in global memory (Wt[k][128], i.e. the transpose of nn.Linear.weight, zero-padded to 128
// columns) streamed through a double-buffered cp.async ring.
#pragma once
#include "common.cuh"

namespace abopt {

constexpr int RT_ROWS = 64;                // rows of the largest tile (R = 8): sizes the staging buffers
constexpr int RT_COLS = 128;
constexpr int RT_THREADS = 256;
constexpr int RT_KC = 32;                  // k-chunk
constexpr int RT_ALD = RT_KC + 4;          // row pitch of a staged A chunk (floats)
constexpr int RT_WLD = RT_COLS + 4;        // row pitch of a staged W chunk
constexpr int RT_ACT_LD = 128 + 4;         // row pitch of a resident 64 x 128 activation tile

struct RowTileSmem {
  float w[2][RT_KC * RT_WLD];              // 2 x 16.5 KB
  float a[2][RT_ROWS * RT_ALD];            // 2 x  9.0 KB
};

template <int R>
__device__ __forceinline__ void rt_zero(float (&acc)[R][4]) {
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
}

// stage W chunk rows [k0, k0+RT_KC) of Wt[K][128] into buffer `buf`
__device__ __forceinline__ void rt_load_w(RowTileSmem& s, int buf, const float* __restrict__ Wt, int k0, int K) {
  // RT_KC x 128 floats = 1024 float4 -> 4 per thread
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int f4 = threadIdx.x + i * RT_THREADS;
    const int k = f4 >> 5, n4 = f4 & 31;
    float* dst = &s.w[buf][k * RT_WLD + n4 * 4];
    if (k0 + k < K) cp_async16(dst, Wt + (size_t)(k0 + k) * RT_COLS + n4 * 4);
    else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// inner product over one staged chunk; A rows come from `arow(r)` = pointer to row r's k-chunk
template <int R, typename ARow>
__device__ __forceinline__ void rt_mma_chunk(float (&acc)[R][4], const float* __restrict__ wbuf, ARow arow, int kc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 2
  for (int k = 0; k < kc; k += 4) {
    float4 w0 = *reinterpret_cast<const float4*>(&wbuf[(k + 0) * RT_WLD + lane * 4]);
    float4 w1 = *reinterpret_cast<const float4*>(&wbuf[(k + 1) * RT_WLD + lane * 4]);
    float4 w2 = *reinterpret_cast<const float4*>(&wbuf[(k + 2) * RT_WLD + lane * 4]);
    float4 w3 = *reinterpret_cast<const float4*>(&wbuf[(k + 3) * RT_WLD + lane * 4]);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(arow(warp * R + r) + k);
      acc[r][0] = fmaf(a.x, w0.x, acc[r][0]); acc[r][1] = fmaf(a.x, w0.y, acc[r][1]);
      acc[r][2] = fmaf(a.x, w0.z, acc[r][2]); acc[r][3] = fmaf(a.x, w0.w, acc[r][3]);
      acc[r][0] = fmaf(a.y, w1.x, acc[r][0]); acc[r][1] = fmaf(a.y, w1.y, acc[r][1]);
      acc[r][2] = fmaf(a.y, w1.z, acc[r][2]); acc[r][3] = fmaf(a.y, w1.w, acc[r][3]);
      acc[r][0] = fmaf(a.z, w2.x, acc[r][0]); acc[r][1] = fmaf(a.z, w2.y, acc[r][1]);
      acc[r][2] = fmaf(a.z, w2.z, acc[r][2]); acc[r][3] = fmaf(a.z, w2.w, acc[r][3]);
      acc[r][0] = fmaf(a.w, w3.x, acc[r][0]); acc[r][1] = fmaf(a.w, w3.y, acc[r][1]);
      acc[r][2] = fmaf(a.w, w3.z, acc[r][2]); acc[r][3] = fmaf(a.w, w3.w, acc[r][3]);
    }
  }
}

// acc += Act(8R x K, resident in smem, pitch lda) * Wt[K][128].   K % 4 == 0.
template <int R>
__device__ __forceinline__ void rt_gemm_smemA(float (&acc)[R][4], RowTileSmem& s, const float* act, int lda,
                                              const float* __restrict__ Wt, int K) {
  const int nchunk = (K + RT_KC - 1) / RT_KC;
  rt_load_w(s, 0, Wt, 0, K);
  cp_async_commit();
  for (int c = 0; c < nchunk; ++c) {
    if (c + 1 < nchunk) rt_load_w(s, (c + 1) & 1, Wt, (c + 1) * RT_KC, K);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int k0 = c * RT_KC;
    const int kc = min(RT_KC, K - k0);
    rt_mma_chunk(acc, s.w[c & 1], [&](int r) { return act + r * lda + k0; }, kc);
    __syncthreads();
  }
}

// acc += A(8R x K gathered from global through `src(row, k)` -> pointer to 4 floats, or nullptr for
// zero fill) * Wt[K][128].   K % RT_KC may be nonzero; K % 4 == 0.
template <int R, typename Src>
__device__ __forceinline__ void rt_gemm_globalA(float (&acc)[R][4], RowTileSmem& s, Src src,
                                                const float* __restrict__ Wt, int K) {
  const int nchunk = (K + RT_KC - 1) / RT_KC;
  auto load_a = [&](int buf, int k0) {
    // 8R rows x 32 k = 64R float4 -> 2 per thread at R = 8
#pragma unroll
    for (int i = 0; i < (64 * R + RT_THREADS - 1) / RT_THREADS; ++i) {
      const int f4 = threadIdx.x + i * RT_THREADS;
      if ((64 * R) % RT_THREADS != 0 && f4 >= 64 * R) break;
      const int r = f4 >> 3, k4 = f4 & 7;
      float* dst = &s.a[buf][r * RT_ALD + k4 * 4];
      const float* g = (k0 + k4 * 4 < K) ? src(r, k0 + k4 * 4) : nullptr;
      if (g) cp_async16(dst, g);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  rt_load_w(s, 0, Wt, 0, K);
  load_a(0, 0);
  cp_async_commit();
  for (int c = 0; c < nchunk; ++c) {
    if (c + 1 < nchunk) { rt_load_w(s, (c + 1) & 1, Wt, (c + 1) * RT_KC, K); load_a((c + 1) & 1, (c + 1) * RT_KC); }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int kc = min(RT_KC, K - c * RT_KC);
    const float* abuf = s.a[c & 1];
    rt_mma_chunk(acc, s.w[c & 1], [&](int r) { return abuf + r * RT_ALD; }, kc);
    __syncthreads();
  }
}

// acc[r][c] += bias[col]
template <int R>
__device__ __forceinline__ void rt_add_bias(float (&acc)[R][4], const float* __restrict__ bias) {
  const int lane = threadIdx.x & 31;
  const float4 b = *reinterpret_cast<const float4*>(bias + lane * 4);
#pragma unroll
  for (int r = 0; r < R; ++r) { acc[r][0] += b.x; acc[r][1] += b.y; acc[r][2] += b.z; acc[r][3] += b.w; }
}
template <int R>
__device__ __forceinline__ void rt_relu(float (&acc)[R][4]) {
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = fmaxf(acc[r][c], 0.f);
}
// write the tile into a resident activation buffer (row-major, pitch lda)
template <int R>
__device__ __forceinline__ void rt_store_act(const float (&acc)[R][4], float* act, int lda) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < R; ++r)
    *reinterpret_cast<float4*>(act + (warp * R + r) * lda + lane * 4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
}
// Reference LayerNorm over the 128 columns of each row (common/layers.py:146-155): biased variance,
// eps inside the sqrt.  In place.
template <int R>
__device__ __forceinline__ void rt_layernorm(float (&v)[R][4], const float* __restrict__ gamma,
                                             const float* __restrict__ beta, float eps) {
  const int lane = threadIdx.x & 31;
  const float4 g = *reinterpret_cast<const float4*>(gamma + lane * 4);
  const float4 b = *reinterpret_cast<const float4*>(beta + lane * 4);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const float mean = warp_sum(v[r][0] + v[r][1] + v[r][2] + v[r][3]) * (1.f / 128.f);
    const float d0 = v[r][0] - mean, d1 = v[r][1] - mean, d2 = v[r][2] - mean, d3 = v[r][3] - mean;
    const float var = warp_sum(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3) * (1.f / 128.f);
    const float sd = sqrtf(var + eps);
    v[r][0] = d0 / sd * g.x + b.x; v[r][1] = d1 / sd * g.y + b.y;
    v[r][2] = d2 / sd * g.z + b.z; v[r][3] = d3 / sd * g.w + b.w;
  }
}

}  // namespace abopt
