// The two kernels that touch the pair tensor z (N, L, L, 64):
//   pair_bias_kernel    z_ij . W_b for every pair and head (ga.py:88-90) -- run ONCE per layer per sampling run: z and the
//                       weights do not change over the T reverse steps (the hoist is in api.cu)
//   pair_stream_kernel  the kernel that streams z once per GABlock: sum_j alpha_ijh z_ijc (ga.py:114-118); alpha comes from
//                       attn_logits_tc_kernel (k_attn_tc.cu)
// Both are persistent, TMA-fed (cp.async.bulk.tensor.2d, 128-byte swizzle, L2 evict-first, mbarrier expect_tx) and do
// their arithmetic as packed FFMA2 (fma.rn.f32x2) on the CUDA cores.
// Why CUDA cores and not tcgen05 here: the contractions are skinny (N = 12 heads) and need fp32-grade accuracy, i.e. a
// 3xTF32 split of z; splitting the 64 KB row block in shared memory plus the operand reads of three MMAs cost ~7 passes
// over the tile (~3600 clk/row of shared-memory bandwidth) against 1536 clk/row of FFMA issue per contraction -- no gain,
// so the tensor cores are kept for the dense GEMMs.
#include <type_traits>
#include "tc.cuh"
#include "params.cuh"
#include "kernels.h"

namespace abopt {

using namespace tc;

constexpr int PS_CONSUMERS = 512;                     // 16 warps (4 per scheduler); thread 0 doubles as the TMA producer
constexpr int PS_THREADS = PS_CONSUMERS;
constexpr int PS_ROWS = PS_CONSUMERS / 2;             // key residues covered per pass of phase A (2 threads per residue)
constexpr int PS_BOX_ROWS = 64;                       // key residues per TMA box
constexpr int PS_HALF_BYTES = PS_BOX_ROWS * 128;      // one TMA box: 64 rows x 32 floats

__device__ __forceinline__ float2 ffma2(float a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  float2 aa = make_float2(a, a);
  ra = *reinterpret_cast<unsigned long long*>(&aa);
  rb = *reinterpret_cast<unsigned long long*>(&b);
  rc = *reinterpret_cast<unsigned long long*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ void consumer_sync() { __syncthreads(); }

__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
               ::"r"(smem_u32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

// byte offset of the 16-byte group q (0..15) of key row j inside a stage (TMA SWIZZLE_128B layout):
// box (j / 64, q / 8) of 64 rows x 128 B; inside a box the group index is XORed with (row & 7)
__device__ __forceinline__ uint32_t zoff(int j, int q) {
  const int jr = j & (PS_BOX_ROWS - 1);
  return (uint32_t)((((j / PS_BOX_ROWS) * 2 + (q >> 3)) * PS_HALF_BYTES) + jr * 128 + (((q & 7) ^ (jr & 7)) << 4));
}

struct PairStreamArgs {
  int L, Lp, b0, nrows;           // nrows = complexes covered by this launch * L; b0 = first complex
  int nstage, stage_bytes, tile_tx_bytes, nbox_rows;   // nbox_rows = 64-row boxes per row-block
  const uint8_t* mask;
  float* alpha;                   // [chunk complex][h][i][Lp]  attention weights from attn_logits_tc_kernel (rows of masked
                                  // queries are zeroed here, ga.py:25)
  float* feat;
  float* feat_lo;                 // tf32 "lo" plane of feat for the out_transform tensor-core GEMM
  float* bias;                    // pair_bias_kernel output, transposed: [complex][h][j][Lp] (query index i contiguous)
};

// ---- TMA producer shared by both kernels: thread 0 walks the CTA's rows (blockIdx.x, + gridDim.x, ...) and loads
//      whole row-blocks z[b, i, :, :] into the ring; `skip_masked` drops rows whose query residue is masked.
struct TileProducer {
  int row, n;
  uint64_t pol;
  __device__ __forceinline__ void issue(const CUtensorMap* zmap, const PairStreamArgs& a, unsigned char* stages, uint64_t* full,
                                        int stride, bool skip_masked) {
    while (row < a.nrows) {
      const int bl = row / a.L, i = row - bl * a.L, b = a.b0 + bl;
      row += stride;
      if (skip_masked && a.mask[(size_t)b * a.L + i] == 0) continue;
      const int s = n % a.nstage;
      ++n;
      unsigned char* st = stages + (size_t)s * a.stage_bytes;
      mbar_expect_tx(&full[s], a.tile_tx_bytes);
      const int grow = (b * a.L + i) * a.L;              // first row of z[b, i] in the (N*L*L, 64) view
      for (int r = 0; r < a.nbox_rows; ++r) {
        tma_load_2d_hint(st + (r * 2 + 0) * PS_HALF_BYTES, zmap, 0, grow + r * PS_BOX_ROWS, &full[s], pol);
        tma_load_2d_hint(st + (r * 2 + 1) * PS_HALF_BYTES, zmap, 32, grow + r * PS_BOX_ROWS, &full[s], pol);
      }
      return;
    }
  }
};

// ------------------------------------------------------------------------------------------ pair bias
// bias(b, h, i, j) = z[b,i,j,:] . W_b[h,:]   (ga.py:88-90).  z and W_b do not change over the T reverse steps, so
// FullDPM.sample runs this ONCE per layer per sampling run (api.cu) instead of once per layer per step.
// thread = (key residue j, 6 of the 12 heads); the head half is warp-uniform so the weight pairs are uniform-register
// operands of FFMA2 (fma.rn.f32x2: scalar z  x  (W[c][h], W[c][h+1])).
template <int JPT>
__global__ void __launch_bounds__(PS_THREADS, 1)
pair_bias_kernel(const __grid_constant__ CUtensorMap zmap, const __grid_constant__ PairBiasPacked pb, const PairStreamArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* stages = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);    // keeps the shared address space
  uint64_t* full = reinterpret_cast<uint64_t*>(stages + (size_t)a.nstage * a.stage_bytes);
  const int L = a.L, Lp = a.Lp;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  TileProducer prod{(int)blockIdx.x, 0, 0};
  if (tid == 0) {
    for (int s = 0; s < a.nstage; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
    tma_prefetch_desc(&zmap);
    prod.pol = policy_evict_first();
    for (int s = 0; s < a.nstage; ++s) prod.issue(&zmap, a, stages, full, gridDim.x, false);
  }
  __syncthreads();
  const int ja = (warp >> 1) * 32 + lane, ch = warp & 1;
  int n = 0;
  for (int row = blockIdx.x; row < a.nrows; row += gridDim.x, ++n) {
    const int bl = row / L, i = row - bl * L, b = a.b0 + bl;
    const int s = n % a.nstage;
    const unsigned char* zs = stages + (size_t)s * a.stage_bytes;
    mbar_wait(&full[s], (n / a.nstage) & 1);
#pragma unroll
    for (int u = 0; u < JPT; ++u) {
      const int j = ja + u * PS_ROWS;
      if (j < L) {
        float2 acc[2][3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { acc[0][k] = make_float2(0.f, 0.f); acc[1][k] = make_float2(0.f, 0.f); }
        auto body = [&](auto chc) {
          constexpr int CH = decltype(chc)::value;
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(zs + zoff(j, q));
            const float zz[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
              for (int k = 0; k < 3; ++k) acc[e & 1][k] = ffma2(zz[e], pb.w[CH][q * 4 + e][k], acc[e & 1][k]);
          }
        };
        if (ch == 0) body(std::integral_constant<int, 0>{}); else body(std::integral_constant<int, 1>{});
        // stored TRANSPOSED, bias[b][h][j][i] (query index contiguous): attn_logits_tc_kernel's epilogue threads are
        // query rows, so a warp reads 32 consecutive i of one key j in one coalesced request.  The scattered 4-byte
        // stores here happen once per sampling run and merge in L2 (neighbouring i are written by neighbouring CTAs).
        float* dst = a.bias + ((size_t)(b * H + ch * (H / 2)) * L + j) * Lp + i;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          dst[(size_t)(2 * k) * L * Lp] = acc[0][k].x + acc[1][k].x;
          dst[(size_t)(2 * k + 1) * L * Lp] = acc[0][k].y + acc[1][k].y;
        }
      }
    }
    fence_async_smem();                                  // order the generic reads before the async-proxy refill
    __syncthreads();
    if (tid == 0) prod.issue(&zmap, a, stages, full, gridDim.x, false);
  }
}

// ------------------------------------------------------------------------------------------ pair aggregation
// pair_stream_kernel: out[b,i,h,c] = sum_j alpha[b,i,j,h] z[b,i,j,c]  (ga.py:114-118), alpha from attn_logits_tc_kernel.
// Persistent CTAs of 8 warps, TWO resident per SM: each owns one shared-memory stage that holds a whole row block
// z[b,i,:,:] (TMA, 128-byte swizzle, L2 evict-first).  While one CTA waits for its next row block the other computes,
// so the SM alternates between them and neither the HBM latency nor the two block barriers per row are exposed.
//   per row: alpha[h][:] (registers, prefetched one row ahead) -> shared [4 residues][head][4]; barrier;
//            lane = 4 channels, half-warp = one 4-residue group, FFMA2 = scalar alpha x channel pair; partial sums ->
//            dedicated scratch; barrier; the stage is refilled at once (thread 0) while 192 threads reduce the 8 slices
//            and store the row of feat (+ its tf32 lo plane).
constexpr int PA_THREADS = 256;
constexpr int PA_SLICES = PA_THREADS / 32;            // 8
constexpr int PA_MAXQ = 6;                            // float4 of alpha per thread per row: 12 heads * L / 4 / 256 <= 6 for L <= 512
constexpr int PA_RED_BYTES = PA_SLICES * H * C * 4;   // 24 KB

__global__ void __launch_bounds__(PA_THREADS, 2)
pair_stream_kernel(const __grid_constant__ CUtensorMap zmap, const PairStreamArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* stages = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);    // keeps the shared address space
  const int L = a.L, Lp = a.Lp;
  const int Lq = (L + 3) & ~3;
  const int nf = Lq / 4;                              // 4-residue groups per row
  float* red = reinterpret_cast<float*>(stages + (size_t)a.nstage * a.stage_bytes);     // [8 slices][12][64] partial sums
  float* als = red + PA_SLICES * H * C;                                                  // [nf][12][4] alpha of the current row
  uint64_t* full = reinterpret_cast<uint64_t*>(als + H * Lq);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  TileProducer prod{(int)blockIdx.x, 0, 0};
  if (tid == 0) {
    for (int s = 0; s < a.nstage; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
    tma_prefetch_desc(&zmap);
    prod.pol = policy_evict_first();
    for (int s = 0; s < a.nstage; ++s) prod.issue(&zmap, a, stages, full, gridDim.x, true);
  }
  __syncthreads();

  auto masked = [&](int row) { const int bl = row / L; return a.mask[(size_t)(a.b0 + bl) * L + (row - bl * L)] == 0; };
  auto next_live = [&](int row) { while (row < a.nrows && masked(row)) row += gridDim.x; return row; };
  // alpha of one query row: 12 * nf float4, thread-strided, prefetched one row ahead into registers
  const int nq = H * nf;
  float4 lg[PA_MAXQ];
  auto load_alpha = [&](int row) {
    if (row < a.nrows) {
      const int bl = row / L, i = row - bl * L;
      const float* base = a.alpha + ((size_t)(bl * H) * L + i) * Lp;
#pragma unroll
      for (int m = 0; m < PA_MAXQ; ++m) {
        const int idx = tid + PA_THREADS * m;
        if (idx < nq) { const int h = idx / nf, f = idx - h * nf; lg[m] = *reinterpret_cast<const float4*>(base + (size_t)h * L * Lp + 4 * f); }
      }
    }
  };
  int live = next_live(blockIdx.x);
  load_alpha(live);

  int n = 0;
  for (int row = blockIdx.x; row < a.nrows; row += gridDim.x) {
    const int bl = row / L, i = row - bl * L, b = a.b0 + bl;
    float* feat_row = a.feat + ((size_t)b * L + i) * NFEAT;
    float* feat_lo_row = a.feat_lo + ((size_t)b * L + i) * NFEAT;
    if (row != live) {
      // masked query: alpha row = 0 (ga.py:25) -> zero pair aggregate
      float* alpha_row0 = a.alpha + ((size_t)(bl * H) * L + i) * Lp;
      for (int o = tid; o < H * C; o += PA_THREADS) { feat_row[o] = 0.f; feat_lo_row[o] = 0.f; }
      for (int o = tid; o < H * Lp; o += PA_THREADS) {
        const int h = o / Lp, j = o - h * Lp;
        alpha_row0[(size_t)h * L * Lp + j] = 0.f;
      }
      continue;
    }
    const int s = n % a.nstage;
    const uint32_t ph = (n / a.nstage) & 1;
    ++n;
    const unsigned char* zs = stages + (size_t)s * a.stage_bytes;

    // ---- alpha -> shared, regrouped as [4 residues][head][4]
#pragma unroll
    for (int m = 0; m < PA_MAXQ; ++m) {
      const int idx = tid + PA_THREADS * m;
      if (idx < nq) { const int h = idx / nf, f = idx - h * nf; reinterpret_cast<float4*>(als)[f * H + h] = lg[m]; }
    }
    live = next_live(row + gridDim.x);
    load_alpha(live);                                    // next row's alpha travels while this row aggregates
    __syncthreads();
    mbar_wait(&full[s], ph);

    // ---- aggregation.  lane = (4 channels = one 16-byte group, half-warp); the two half-warps of a warp take different
    //      4-residue groups, so one iteration issues 4 LDS.128 of z + 12 LDS.128 of alpha (2 addresses each) for 96 FFMA2
    //      (scalar alpha x channel pair) -- the shared-memory instruction rate, not its bandwidth, is what limits this loop
    float2 acc[H][2];
#pragma unroll
    for (int h = 0; h < H; ++h) { acc[h][0] = make_float2(0.f, 0.f); acc[h][1] = make_float2(0.f, 0.f); }
    {
      // Row j0 + k (j0 % 4 == 0) stores the 16-byte group q at ((q ^ k) ^ (j0 & 4)) * 16 inside its box half -- see zoff().
      const int c4 = lane & 15, q = c4 & 7;
      const uint32_t lane_off = (uint32_t)((c4 >> 3) * PS_HALF_BYTES);
      uint32_t xk[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) xk[k] = lane_off + k * 128 + ((q ^ k) << 4);
      for (int g = warp * 2 + (lane >> 4); g < nf; g += 2 * PA_SLICES) {
        const int j0 = 4 * g;
        const unsigned char* zr = zs + (j0 / PS_BOX_ROWS) * (2 * PS_HALF_BYTES) + (j0 & (PS_BOX_ROWS - 1)) * 128;
        const uint32_t flip = (uint32_t)(j0 & 4) << 4;
        float4 zv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) zv[k] = *reinterpret_cast<const float4*>(zr + (xk[k] ^ flip));
        const float4* ap = reinterpret_cast<const float4*>(als) + g * H;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float4 av = ap[h];
          acc[h][0] = ffma2(av.x, make_float2(zv[0].x, zv[0].y), acc[h][0]); acc[h][1] = ffma2(av.x, make_float2(zv[0].z, zv[0].w), acc[h][1]);
          acc[h][0] = ffma2(av.y, make_float2(zv[1].x, zv[1].y), acc[h][0]); acc[h][1] = ffma2(av.y, make_float2(zv[1].z, zv[1].w), acc[h][1]);
          acc[h][0] = ffma2(av.z, make_float2(zv[2].x, zv[2].y), acc[h][0]); acc[h][1] = ffma2(av.z, make_float2(zv[2].z, zv[2].w), acc[h][1]);
          acc[h][0] = ffma2(av.w, make_float2(zv[3].x, zv[3].y), acc[h][0]); acc[h][1] = ffma2(av.w, make_float2(zv[3].z, zv[3].w), acc[h][1]);
        }
      }
      // the two half-warps hold partial sums of the same channels: combine, lanes 0..15 write the warp's slice
#pragma unroll
      for (int h = 0; h < H; ++h) {
        acc[h][0].x += __shfl_xor_sync(0xffffffffu, acc[h][0].x, 16); acc[h][0].y += __shfl_xor_sync(0xffffffffu, acc[h][0].y, 16);
        acc[h][1].x += __shfl_xor_sync(0xffffffffu, acc[h][1].x, 16); acc[h][1].y += __shfl_xor_sync(0xffffffffu, acc[h][1].y, 16);
      }
      if (lane < 16)
#pragma unroll
        for (int h = 0; h < H; ++h)
          *reinterpret_cast<float4*>(red + warp * (H * C) + h * C + c4 * 4) = make_float4(acc[h][0].x, acc[h][0].y, acc[h][1].x, acc[h][1].y);
    }
    fence_async_smem();                                  // generic reads of the stage ordered before the async-proxy refill
    __syncthreads();                                     // every read of the stage is done, every partial sum is visible
    if (tid == 0) prod.issue(&zmap, a, stages, full, gridDim.x, true);       // refill at once: the load overlaps the rest
    if (tid < H * C / 4) {
      float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < PA_SLICES; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(red + k * (H * C) + tid * 4);
        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
      }
      *reinterpret_cast<float4*>(feat_row + tid * 4) = sum;
      *reinterpret_cast<float4*>(feat_lo_row + tid * 4) = make_float4(tf32_lo(sum.x), tf32_lo(sum.y), tf32_lo(sum.z), tf32_lo(sum.w));
    }
    // the next row's first barrier (alpha staged) also separates this reduction from the next partial-sum writes
  }
}

// ------------------------------------------------------------------------------------------ host side
static int g_sm_count = 0;

cudaError_t pair_stream_init() {
  cudaError_t e;
  int dev = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
  const int mx = 227 * 1024;
  if ((e = cudaFuncSetAttribute(pair_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(pair_bias_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(pair_bias_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(pair_bias_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)) != cudaSuccess) return e;
  return cudaSuccess;
}

// z viewed as a 2-D fp32 matrix [(N*L*L) rows][64]; boxes of [<=64 rows][32 floats], 128-byte swizzle
bool make_pair_tmap(CUtensorMap* m, const float* z, size_t total_rows, int* box_rows_out) {
  const uint32_t box_rows = total_rows < (size_t)PS_BOX_ROWS ? (uint32_t)total_rows : (uint32_t)PS_BOX_ROWS;
  *box_rows_out = (int)box_rows;
  return make_tmap_2d(m, z, total_rows, C, C, box_rows, 32);
}

static bool fill_args(PairStreamArgs& a, int nb, int b0, int L, int Lp, int box_rows, size_t fixed, size_t* smem) {
  a.L = L; a.Lp = Lp; a.b0 = b0; a.nrows = nb * L;
  a.nbox_rows = (L + PS_BOX_ROWS - 1) / PS_BOX_ROWS;
  a.stage_bytes = a.nbox_rows * 2 * PS_HALF_BYTES;
  a.tile_tx_bytes = a.nbox_rows * 2 * box_rows * 128;
  int nstage = (int)((227 * 1024 - fixed) / a.stage_bytes);
  if (nstage < 1) return false;
  if (nstage > 4) nstage = 4;
  a.nstage = nstage;
  *smem = (size_t)nstage * a.stage_bytes + fixed;
  return true;
}

// bias[b][h][i][Lp] for complexes [b0, b0 + nb)
bool launch_pair_bias(int nb, int b0, int L, int Lp, const CUtensorMap& zmap, int box_rows, const PairBiasPacked& pb, float* bias,
                      cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  PairStreamArgs a{};
  size_t smem = 0;
  if (!fill_args(a, nb, b0, L, Lp, box_rows, 8 * 8 + 1024, &smem)) return false;
  a.bias = bias;
  int grid = g_sm_count > 0 ? g_sm_count : 148;
  if (grid > a.nrows) grid = a.nrows;
  if (L <= PS_ROWS) pair_bias_kernel<1><<<grid, PS_THREADS, smem, st>>>(zmap, pb, a);
  else if (L <= 2 * PS_ROWS) pair_bias_kernel<2><<<grid, PS_THREADS, smem, st>>>(zmap, pb, a);
  else pair_bias_kernel<3><<<grid, PS_THREADS, smem, st>>>(zmap, pb, a);
  return true;
}

bool launch_pair_stream(int nb, int b0, int L, int Lp, const CUtensorMap& zmap, int box_rows, const uint8_t* mask,
                        float* alpha, float* feat, float* feat_lo, cudaStream_t st) {
  ProfScope prof__(KK_PAIR, st);
  if (H * ((L + 3) / 4) > PA_THREADS * PA_MAXQ) return false;
  PairStreamArgs a{};
  const int Lq = (L + 3) & ~3;
  a.L = L; a.Lp = Lp; a.b0 = b0; a.nrows = nb * L;
  a.nbox_rows = (L + PS_BOX_ROWS - 1) / PS_BOX_ROWS;
  a.stage_bytes = a.nbox_rows * 2 * PS_HALF_BYTES;
  a.tile_tx_bytes = a.nbox_rows * 2 * box_rows * 128;
  const size_t fixed = (size_t)PA_RED_BYTES + (size_t)H * Lq * 4 + 8 * 8 + 1024;
  // two CTAs per SM when a stage + the fixed part fit half of the shared memory, else one CTA with what fits
  const size_t half = (227 * 1024) / 2 - 1024;
  int nstage, per_sm;
  if (a.stage_bytes + fixed <= half) { per_sm = 2; nstage = (int)((half - fixed) / a.stage_bytes); }
  else { per_sm = 1; nstage = (int)((227 * 1024 - fixed) / a.stage_bytes); }
  if (nstage < 1) return false;
  if (nstage > 4) nstage = 4;
  a.nstage = nstage;
  a.mask = mask; a.alpha = alpha; a.feat = feat; a.feat_lo = feat_lo;
  const size_t smem = (size_t)nstage * a.stage_bytes + fixed;
  int grid = (g_sm_count > 0 ? g_sm_count : 148) * per_sm;
  if (grid > a.nrows) grid = a.nrows;
  pair_stream_kernel<<<grid, PA_THREADS, smem, st>>>(zmap, a);
  return true;
}

}  // namespace abopt
