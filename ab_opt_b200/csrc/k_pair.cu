// pair_stream_kernel: the kernel that streams the pair tensor z (N, L, L, 64) ONCE per GABlock.
// For every query residue (b, i) it consumes the row-block z[b, i, :, :] (L x 64 fp32) and produces
//   pair bias        z_ij . W_b                                   ga.py:88-90
//   attention        softmax_j((node + spatial + pair) * sqrt(1/3)), masked          ga.py:11-26,166
//   pair aggregate   sum_j alpha_ijh z_ijc                          ga.py:114-118
// The node + spatial logits S arrive from logits_kernel through L2; alpha leaves through L2 for aggr_kernel.
//
// Structure: persistent CTAs (one per SM), 16 warps.
//   producer : thread 0 walks the CTA's rows and issues TMA tensor loads (cp.async.bulk.tensor.2d, 128-byte
//              swizzle, L2 evict-first) of whole row-blocks into a ring of shared-memory stages, completion on
//              mbarriers (expect_tx).  A stage is refilled right after the first block barrier of the NEXT row,
//              i.e. as soon as every warp is known to have left it -- no "empty" barriers, no extra warp.
//   consumers: phase A  thread = (key residue j, 6 of the 12 heads): 64 channels x 3 head pairs from the swizzled
//                       256 B row, issued as packed FFMA2 (fma.rn.f32x2: scalar z  x  (W[c][h], W[c][h+1]) pairs
//                       that live in the constant bank -> uniform registers), then scale / mask;
//              phase B  block softmax over j through a [12][L] shared tile (one warp per head);
//              phase C  lane = channel pair, warp = 1 of 16 j-slices: 12 x 2 accumulators of alpha x z, FFMA2
//                       again (scalar alpha x channel pair), cross-slice reduction through the drained stage.
// Why CUDA cores and not tcgen05 here: the two contractions are skinny (N = 12 heads) and need fp32-grade
// accuracy, i.e. a 3xTF32 split of z; splitting z in shared memory plus the operand reads of the three MMAs
// cost ~7 passes over the 64 KB tile (~3600 clk/row of shared-memory bandwidth) against 3072 clk/row of
// FFMA2 issue -- no gain, so the tensor cores are kept for the node-feature linears.
#include <type_traits>
#include "tc.cuh"
#include "params.cuh"
#include "kernels.h"

namespace abopt {

using namespace tc;

constexpr int PS_CONSUMERS = 512;                     // 16 warps (4 per scheduler); thread 0 doubles as the TMA producer
constexpr int PS_THREADS = PS_CONSUMERS;
constexpr int PS_ROWS = PS_CONSUMERS / 2;             // key residues covered per pass of phase A (2 threads per residue)
constexpr int PS_SLICES = PS_CONSUMERS / 32;          // j-slices in phase C (one per warp)
constexpr int PS_BOX_ROWS = 64;                       // key residues per TMA box
constexpr int PS_HALF_BYTES = PS_BOX_ROWS * 128;      // one TMA box: 64 rows x 32 floats
constexpr int PS_RED_BYTES = PS_SLICES * H * C * 4;   // cross-slice reduction scratch (lives in the drained stage)

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

// ------------------------------------------------------------------------------------------ softmax + pair aggregation
constexpr int PS_MAXF = 4;        // float4 groups of an alpha row per lane: L <= 32 * 4 * 4 = 512

__global__ void __launch_bounds__(PS_THREADS, 1)
pair_stream_kernel(const __grid_constant__ CUtensorMap zmap, const PairStreamArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* stages = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);    // keeps the shared address space
  const int L = a.L, Lp = a.Lp;
  const int Lq = (L + 3) & ~3;                       // pitch of the attention tile (floats), == Lp
  const int nf = Lq / 4;
  float* als = reinterpret_cast<float*>(stages + (size_t)a.nstage * a.stage_bytes);    // [Lq / 4][12][4] alpha of the current row
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
  // alpha of one query row, one warp per head, lane-strided float4 groups (prefetched one row ahead)
  float4 lg[PS_MAXF];
  auto load_logits = [&](int row) {
    if (warp < H && row < a.nrows) {
      const int bl = row / L, i = row - bl * L;
      const float4* src = reinterpret_cast<const float4*>(a.alpha + ((size_t)(bl * H + warp) * L + i) * Lp);
#pragma unroll
      for (int m = 0; m < PS_MAXF; ++m) { const int f = lane + 32 * m; if (f < nf) lg[m] = src[f]; }
    }
  };
  int live = next_live(blockIdx.x);
  load_logits(live);

  int n = 0;
  for (int row = blockIdx.x; row < a.nrows; row += gridDim.x) {
    const int bl = row / L, i = row - bl * L, b = a.b0 + bl;
    float* feat_row = a.feat + ((size_t)b * L + i) * NFEAT;
    float* feat_lo_row = a.feat_lo + ((size_t)b * L + i) * NFEAT;
    if (row != live) {
      // masked query: alpha row = 0 (ga.py:25) -> zero pair aggregate
      float* alpha_row0 = a.alpha + ((size_t)(bl * H) * L + i) * Lp;
      for (int o = tid; o < H * C; o += PS_THREADS) { feat_row[o] = 0.f; feat_lo_row[o] = 0.f; }
      for (int o = tid; o < H * Lp; o += PS_THREADS) {
        const int h = o / Lp, j = o - h * Lp;
        alpha_row0[(size_t)h * L * Lp + j] = 0.f;
      }
      continue;
    }
    const int s = n % a.nstage;
    const uint32_t ph = (n / a.nstage) & 1;
    ++n;
    const unsigned char* zs = stages + (size_t)s * a.stage_bytes;

    // ---- alpha[h][:] of this query row (prefetched registers) -> shared, regrouped as [4 residues][head][4]
    if (warp < H) {
      float4* dsts = reinterpret_cast<float4*>(als) + warp;
#pragma unroll
      for (int k = 0; k < PS_MAXF; ++k) {
        const int f = lane + 32 * k;
        if (f < nf) dsts[f * H] = lg[k];
      }
    }
    live = next_live(row + gridDim.x);
    load_logits(live);                                   // next row's logits travel while this row aggregates
    __syncthreads();
    // every thread has left the previous row: its stage is drained -> thread 0 refills it (the generic-proxy accesses
    // to that stage were ordered before this async-proxy write by the fence each thread issued at the end of the row)
    if (tid == 0 && a.nstage > 1 && n > 1) prod.issue(&zmap, a, stages, full, gridDim.x, true);
    mbar_wait(&full[s], ph);

    // ---- pair aggregation out[h][c] = sum_j alpha[j][h] z[j][c]   (ga.py:114-118)
    //      lane = channel pair (a warp reads whole 256 B rows), warp = slice of 4-residue groups; alpha[h][j..j+3]
    //      is one broadcast LDS.128; FFMA2 = scalar alpha x channel pair
    float2 acc[H];
#pragma unroll
    for (int h = 0; h < H; ++h) acc[h] = make_float2(0.f, 0.f);
    {
      // lane -> 8 bytes of the 256 B row: box half (lane >> 4), 16-byte group q = (lane >> 1) & 7, low / high 8 bytes.
      // Row j0 + k (j0 % 4 == 0) stores group q at ((q ^ k) ^ (j0 & 4)) * 16 -- see zoff().
      const int q = (lane >> 1) & 7;
      const uint32_t lane_off = (uint32_t)((lane >> 4) * PS_HALF_BYTES + (lane & 1) * 8);
      uint32_t xk[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) xk[k] = lane_off + k * 128 + ((q ^ k) << 4);
      for (int g = warp; g < nf; g += PS_SLICES) {
        const int j0 = 4 * g;
        const unsigned char* zr = zs + (j0 / PS_BOX_ROWS) * (2 * PS_HALF_BYTES) + (j0 & (PS_BOX_ROWS - 1)) * 128;
        const uint32_t flip = (uint32_t)(j0 & 4) << 4;
        float2 zv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) zv[k] = *reinterpret_cast<const float2*>(zr + (xk[k] ^ flip));
        const float4* ap = reinterpret_cast<const float4*>(als) + g * H;
        float4 av[H];
#pragma unroll
        for (int h = 0; h < H; ++h) av[h] = ap[h];
#pragma unroll
        for (int h = 0; h < H; ++h) acc[h] = ffma2(av[h].x, zv[0], acc[h]);
#pragma unroll
        for (int h = 0; h < H; ++h) acc[h] = ffma2(av[h].y, zv[1], acc[h]);
#pragma unroll
        for (int h = 0; h < H; ++h) acc[h] = ffma2(av[h].z, zv[2], acc[h]);
#pragma unroll
        for (int h = 0; h < H; ++h) acc[h] = ffma2(av[h].w, zv[3], acc[h]);
      }
    }
    __syncthreads();                                     // every read of the stage is done -> reuse it as scratch
    float* red = reinterpret_cast<float*>(const_cast<unsigned char*>(zs));
#pragma unroll
    for (int h = 0; h < H; ++h) *reinterpret_cast<float2*>(red + warp * (H * C) + h * C + lane * 2) = acc[h];
    __syncthreads();
    if (tid < H * C / 4) {
      float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < PS_SLICES; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(red + k * (H * C) + tid * 4);
        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
      }
      *reinterpret_cast<float4*>(feat_row + tid * 4) = sum;
      *reinterpret_cast<float4*>(feat_lo_row + tid * 4) = make_float4(tf32_lo(sum.x), tf32_lo(sum.y), tf32_lo(sum.z), tf32_lo(sum.w));
    }
    // generic-proxy accesses to this stage must be ordered before the next TMA (async proxy) write into it
    fence_async_smem();
    if (a.nstage == 1) { __syncthreads(); if (tid == 0) prod.issue(&zmap, a, stages, full, gridDim.x, true); }
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
  const int tile_bytes = a.nbox_rows * 2 * PS_HALF_BYTES;
  a.stage_bytes = tile_bytes > PS_RED_BYTES ? tile_bytes : PS_RED_BYTES;
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
  if (L > 32 * 4 * PS_MAXF) return false;
  PairStreamArgs a{};
  size_t smem = 0;
  const int Lq = (L + 3) & ~3;
  if (!fill_args(a, nb, b0, L, Lp, box_rows, (size_t)H * Lq * 4 + 8 * 8 + 1024, &smem)) return false;
  a.mask = mask; a.alpha = alpha; a.feat = feat; a.feat_lo = feat_lo;
  int grid = g_sm_count > 0 ? g_sm_count : 148;
  if (grid > a.nrows) grid = a.nrows;
  pair_stream_kernel<<<grid, PS_THREADS, smem, st>>>(zmap, a);
  return true;
}

}  // namespace abopt
