// The two kernels that touch the pair tensor z (N, L, L, 64):
//   pair_bias_kernel    z_ij . W_b for every pair and head (ga.py:88-90) -- run ONCE per layer per sampling run: z and the
//                       weights do not change over the T reverse steps (the hoist is in api.cu)
//   pair_stream_kernel  the kernel that streams z once per GABlock: sum_j alpha_ijh z_ijc (ga.py:114-118); alpha comes from
//                       attn_logits_persist_kernel (k_attn_tc.cu)
// pair_bias_kernel is barrier-free (every warp owns whole query rows and streams its row block through a private TMA ring,
// packed FFMA2 on the CUDA cores: it runs 6 times per sampling run).  pair_stream_kernel, which runs 600 times per run, puts
// the contraction on tcgen05 with z as the TMEM-resident A operand; ctx_delta_kernel is the context-cache companion of the
// first GABlock (see each kernel's header).
#include <cstdlib>
#include <type_traits>
#include "tc.cuh"
#include "params.cuh"
#include "kernels.h"

namespace abopt {

using namespace tc;

__device__ __forceinline__ float2 ffma2(float a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  float2 aa = make_float2(a, a);
  ra = *reinterpret_cast<unsigned long long*>(&aa);
  rb = *reinterpret_cast<unsigned long long*>(&b);
  rc = *reinterpret_cast<unsigned long long*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
               ::"r"(smem_u32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

// ------------------------------------------------------------------------------------------ pair bias
// pair_bias_kernel: bias[b,h,i,j] = z[b,i,j,:] . W_b[h,:]   (ga.py:88-90), stored like alpha: [b][h][i][Lp], key index
// contiguous.  z and W_b do not change over the T reverse steps, so FullDPM.sample runs this ONCE per layer per sampling run
// (api.cu); a training step runs it once per layer.
// Barrier-free: 12 independent warps per SM, each owns whole query rows (b, i) and
// streams the row block z[b,i,:,:] through a private 2-stage ring of 32-key chunks (two TMA boxes of [32 keys][32 channels],
// 128-byte swizzle, L2 evict-first).  lane = key residue: 16 conflict-free LDS.128 of its z row feed 384 FFMA2
// (scalar z  x  (W[c][h], W[c][h+1]) head pairs that live in the constant bank -> uniform registers); the 12 results per lane
// leave as 12 coalesced 128-byte stores.  Columns L <= j < Lp are written as zeros (the logits kernel reads whole chunks).
constexpr int PB_WARPS = 12, PB_THREADS = PB_WARPS * 32, PB_STAGES = 2, PB_CJ = 32;
constexpr int PB_BOX_BYTES = PB_CJ * 128;               // one TMA box: 32 keys x 32 channels
constexpr int PB_STAGE_BYTES = 2 * PB_BOX_BYTES;        // 8 KB: channels 0..31 | 32..63
constexpr int PB_WARP_BYTES = PB_STAGES * PB_STAGE_BYTES;
constexpr int PB_SMEM = PB_WARPS * PB_WARP_BYTES + PB_WARPS * PB_STAGES * 8 + 1024;

struct PairBiasArgs {
  int L, Lp, b0, nrows, nchunk;   // nrows = complexes covered by this launch * L; nchunk = ceil(Lp / 32)
  int box_bytes;                  // bytes one TMA box delivers (32 keys x 128 B, fewer rows when N*L*L < 32)
  float* bias;                    // [complex][h][i][Lp]
};

__global__ void __launch_bounds__(PB_THREADS, 1)
pair_bias_kernel(const __grid_constant__ CUtensorMap zmap, const __grid_constant__ PairBiasPacked pb, const PairBiasArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = a.L, Lp = a.Lp;
  unsigned char* wst = base + warp * PB_WARP_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(base + PB_WARPS * PB_WARP_BYTES) + warp * PB_STAGES;
  if (lane == 0) {
    for (int s = 0; s < PB_STAGES; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
    tma_prefetch_desc(&zmap);
  }
  __syncwarp();
  const int stride = gridDim.x * PB_WARPS;
  const int first = warp * gridDim.x + blockIdx.x;
  const int dbl = stride / L, di = stride - dbl * L;
  auto advance = [&](int& bl, int& i) { bl += dbl; i += di; if (i >= L) { i -= L; ++bl; } };

  // producer cursor (lane 0 issues): chunk by chunk, PB_STAGES chunks ahead of the consumer
  int prow = first, pbl = first / L, pi = first - pbl * L, pjc = 0, ps = 0;
  const uint64_t pol = policy_evict_first();
  auto issue = [&]() {
    if (prow >= a.nrows) return;
    if (lane == 0) {
      unsigned char* st = wst + ps * PB_STAGE_BYTES;
      const int grow = ((a.b0 + pbl) * L + pi) * L + pjc * PB_CJ;      // first row of the chunk in the (N*L*L, 64) view
      mbar_expect_tx(&full[ps], 2 * a.box_bytes);                       // rows past the end of the tensor arrive as zeros
      tma_load_2d_hint(st, &zmap, 0, grow, &full[ps], pol);
      tma_load_2d_hint(st + PB_BOX_BYTES, &zmap, 32, grow, &full[ps], pol);
    }
    if (++ps == PB_STAGES) ps = 0;
    if (++pjc == a.nchunk) { pjc = 0; prow += stride; advance(pbl, pi); }
  };
  for (int s = 0; s < PB_STAGES; ++s) issue();

  int cs = 0;
  uint32_t cph = 0;
  int bl = first / L, i = first - bl * L;
  for (int row = first; row < a.nrows; row += stride, advance(bl, i)) {
    const int b = a.b0 + bl;
    float* out = a.bias + ((size_t)(b * H) * L + i) * Lp;               // + h * L * Lp + j
    for (int jc = 0; jc < a.nchunk; ++jc) {
      mbar_wait(&full[cs], cph);
      const unsigned char* zr = wst + cs * PB_STAGE_BYTES + lane * 128;  // this lane's key row; 16-byte chunk q sits at q ^ (lane & 7)
      float2 acc[2][3];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf)
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[hf][k] = make_float2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(zr + (q >> 3) * PB_BOX_BYTES + (((q & 7) ^ (lane & 7)) << 4));
        const float zz[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf)
#pragma unroll
            for (int k = 0; k < 3; ++k) acc[hf][k] = ffma2(zz[e], pb.w[hf][q * 4 + e][k], acc[hf][k]);
      }
      __syncwarp();                                      // every lane's reads of the stage have completed -> refill it
      issue();
      if (++cs == PB_STAGES) { cs = 0; cph ^= 1u; }
      const int j = jc * PB_CJ + lane;
      if (j < Lp) {
        const bool in = j < L;                           // padding columns hold zeros
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            out[(size_t)(hf * 6 + 2 * k) * L * Lp + j] = in ? acc[hf][k].x : 0.f;
            out[(size_t)(hf * 6 + 2 * k + 1) * L * Lp + j] = in ? acc[hf][k].y : 0.f;
          }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ pair aggregation
// out[b,i,h,c] = sum_j alpha[b,i,j,h] z[b,i,j,c]  (ga.py:114-118), alpha from attn_logits_persist_kernel.
// The rows to visit come from a ROW LIST (pair_rows_build_kernel: live rows from the front, masked-but-needed rows from the
// back), so the kernel evaluates no masks, focus lists or integer divisions in its loops.
struct PairRowsArgs {
  int L, Lp, b0, nrows, nchunk;   // nrows = complexes covered by this launch * L; b0 = first complex; nchunk = ceil(L / 32)
  const float* z;                 // (N, L, L, 64)
  float* alpha;                   // [chunk complex][h][i][Lp]; rows of masked queries are zeroed here (ga.py:25)
  float* feat; int feat_ld;       // output rows of H * C floats, row pitch feat_ld (the 1824-wide feature rows, or a 768-wide cache)
  const int4* list;               // [nrows] (complex within the launch, residue, output row, -)
  const int* count;               // [2] live rows (list front) | masked rows that still need their zeros (list back)
};

// Row list of one launch, in row order: a row is NEEDED when the caller consumes it (focus mode: cidx[row] >= 0 = its compact
// output row; otherwise every row), LIVE when it is needed and its query residue is unmasked.  One CTA; runs once per sampling
// run (the masks are loop invariants) or once per stand-alone block call.
__global__ void __launch_bounds__(1024)
pair_rows_build_kernel(int nrows, int L, int b0, const uint8_t* __restrict__ mask, const int* __restrict__ cidx, int compact,
                       int4* __restrict__ list, int* __restrict__ count) {
  __shared__ int wl[32], wd[32], base[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { base[0] = 0; base[1] = 0; }
  __syncthreads();
  for (int r0 = 0; r0 < nrows; r0 += 1024) {
    const int r = r0 + tid;
    bool live = false, dead = false;
    int bl = 0, i = 0, orow = 0;
    if (r < nrows) {
      bl = r / L; i = r - bl * L;
      const size_t gr = (size_t)(b0 + bl) * L + i;
      orow = cidx ? cidx[gr] : (int)gr;
      live = orow >= 0 && mask[gr] != 0;
      dead = orow >= 0 && !live;
      if (!compact) orow = (int)gr;               // the listed rows keep their place in the full-size output
    }
    const unsigned ml = __ballot_sync(0xffffffffu, live), md = __ballot_sync(0xffffffffu, dead);
    if (lane == 0) { wl[warp] = __popc(ml); wd[warp] = __popc(md); }
    __syncthreads();
    int pl = base[0], pd = base[1];
    for (int w = 0; w < warp; ++w) { pl += wl[w]; pd += wd[w]; }
    const unsigned lt = (1u << lane) - 1u;
    if (live) list[pl + __popc(ml & lt)] = make_int4(bl, i, orow, 0);
    if (dead) list[nrows - 1 - (pd + __popc(md & lt))] = make_int4(bl, i, orow, 0);
    __syncthreads();
    if (tid == 0) {
      int sl = 0, sd = 0;
      for (int w = 0; w < 32; ++w) { sl += wl[w]; sd += wd[w]; }
      base[0] += sl; base[1] += sd;
    }
    __syncthreads();
  }
  if (tid == 0) { count[0] = base[0]; count[1] = base[1]; }
}

// pair_stream_kernel: the contraction on the tensor cores (3xTF32), the z row blocks as the A operand read from TENSOR MEMORY
// (TS-mode MMAs).  One persistent CTA per SM; a tile is TWO query rows of the row list:
//     D[m = row * 64 + channel][n = row' * 16 + head] = sum_j z[row][j][channel] * alpha[row'][head][j]
// (M = 128, N = 32, the two diagonal blocks are the result, the off-diagonal ones are never read), K = the keys in chunks of 32.
//   warp 0        producer: per chunk two 8 KB bulk copies of z (L2 evict-first) and two alpha boxes [12 heads][32 keys]
//                 (128-byte swizzle = the K-major B operand as it lands; heads 12..15 of a row stay zero) through an 8-stage ring
//   warps 6-17    three transposer groups of four warps, chunk g goes to group g % 3: thread = TMEM lane = (row, channel)
//                 reads its column of the chunk from shared memory (32 conflict-free scalar loads), and stores it into one of
//                 four TMEM operand slots as raw | tf32-lo planes (the tensor core ignores the low 13 mantissa bits of the raw
//                 plane); the group also builds the lo plane of the alpha tile
//   warp 1        MMA issuer: per chunk 4 k-steps x (hi*hi, hi*lo, lo*hi) of 128 x 32 x 8 -- hi*hi into one of two main
//                 accumulators (one per half of the keys), the corrections into a third (short chains, see k_tc.cu)
//   warps 2-5     epilogue: thread = (row, channel) adds the three 12-head sums and stores them (a warp writes 128
//                 contiguous bytes per head); also zeroes the masked rows of the list
// Round 1 and most of round 2 ran this contraction on the CUDA cores (12 barrier-free warps per SM, per float4 of z one LDS.128 +
// 24 FFMA2 into 96 live accumulators, 168 registers, 3 warps per scheduler): 242-249 us per full C2 launch = 0.67-0.70 of the HBM
// peak, bound by FFMA2 issue latency, not by memory.  Here a float4 of z costs 4 LDS + 4 LOP + 4 FADD and 1/4 of a
// tcgen05.st, nothing stays in registers between chunks, and the tensor-pipe floor (12 MMAs x 16 clk per chunk = 192 clk against
// ~840 clk of HBM time per chunk) is far below the memory time: 207 us, i.e. z + alpha + the output rows move at 0.99 of the
// measured copy bandwidth.  Ring depth 8 measured best (9: 221 us); L2-sized batch chunks lose (2 x 109 / 4 x 57 us).
// Tiles are walked from the LAST list entry to the first: the logits kernel has just written alpha in forward order, so the tail
// of it is what the 126 MB L2 still holds (211 -> 207 us); aggr_persist_kernel then walks forward for the same reason.
constexpr int PX_CJ = 32;                                 // keys per chunk
constexpr int PX_ZROW = PX_CJ * C * 4;                    // 8192: one query row's chunk of z
constexpr int PX_AB = 32 * 128;                           // 4096: alpha tile [2 rows x 16 heads][32 keys]
constexpr int PX_STAGE = 2 * PX_ZROW + 2 * PX_AB;         // z row 0 | z row 1 | alpha raw | alpha lo
constexpr int PX_TX_A = H * PX_CJ * 4;                    // 1536: what one alpha box delivers
#ifndef ABOPT_PX_NST
#define ABOPT_PX_NST 8
#endif
constexpr int PX_NST = ABOPT_PX_NST, PX_NSLOT = 4, PX_NG = 3;
#define PX_TILE(t) (ntiles - 1 - (t))
constexpr int PX_TW0 = 6;                                 // first transposer warp
constexpr int PX_THREADS = (PX_TW0 + 4 * PX_NG) * 32;     // 576
constexpr int PX_BAR_OFF = PX_NST * PX_STAGE;
constexpr int PX_SMEM = PX_BAR_OFF + 512 + 1024;
constexpr uint32_t PX_TM_A = 0, PX_TM_ACC = 256, PX_ACC_COLS = 96;      // TMEM: 4 x (raw 32 | lo 32) | 2 x (main 32 | main 32 | corrections 32)
static_assert(PX_SMEM <= 227 * 1024, "pair_stream_kernel: shared memory");
static_assert((3 * PX_NST + PX_NSLOT + 4) * 8 + 4 <= 512, "pair_stream_kernel: barrier block");

__device__ __forceinline__ void px_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void px_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
        "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
        "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
        "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}

__global__ void __launch_bounds__(PX_THREADS, 1)
pair_stream_kernel(const __grid_constant__ CUtensorMap amap, const PairRowsArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + PX_BAR_OFF);
  uint64_t* ready = full + PX_NST;           // z chunk in its TMEM slot, alpha lo plane built
  uint64_t* empty = ready + PX_NST;          // the MMAs reading the stage have completed
  uint64_t* ta_free = empty + PX_NST;        // [PX_NSLOT] the MMAs reading the TMEM operand slot have completed
  uint64_t* acc_full = ta_free + PX_NSLOT;   // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = a.L, nkb = a.nchunk, gsz = (nkb + 1) / 2;
  const int nlive = a.count[0], ndead = a.count[1];
  const int ntiles = (nlive + 1) >> 1;

  // a short last chunk leaves the rest of its z slot untouched, a single-row tile its second half, and heads 12..15 of the
  // alpha tile are never written: start from zeros so that whatever is stale is finite (it meets alpha = 0 or unused columns)
  for (int o = threadIdx.x; o < PX_NST * PX_STAGE / 16; o += PX_THREADS) reinterpret_cast<float4*>(smem)[o] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (threadIdx.x == 0) {
    for (int s = 0; s < PX_NST; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], 4); mbar_init(&empty[s], 1); }
    for (int s = 0; s < PX_NSLOT; ++s) mbar_init(&ta_free[s], 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
    mbar_fence_init();
    tma_prefetch_desc(&amap);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      const uint64_t pol = policy_evict_first();
      int g = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int pt = PX_TILE(tile);
        const int4 e0 = a.list[2 * pt];
        const bool two = 2 * pt + 1 < nlive;
        const int4 e1 = two ? a.list[2 * pt + 1] : e0;
        const float* z0 = a.z + ((size_t)(a.b0 + e0.x) * L + e0.y) * L * C;
        const float* z1 = a.z + ((size_t)(a.b0 + e1.x) * L + e1.y) * L * C;
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int s = g % PX_NST;
          mbar_wait(&empty[s], ((g / PX_NST) & 1) ^ 1);
          const uint32_t st = smem_u32(smem + s * PX_STAGE), bar = smem_u32(&full[s]);
          const int j0 = kb * PX_CJ;
          const int nj = (L - j0 < PX_CJ) ? (L - j0) : PX_CJ;
          const uint32_t zb = (uint32_t)(nj * C * 4);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((two ? 2u : 1u) * (zb + PX_TX_A)) : "memory");
          asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                       ::"r"(st), "l"(z0 + (size_t)j0 * C), "r"(zb), "r"(bar), "l"(pol) : "memory");
          asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                       ::"r"(st + 2 * PX_ZROW), "l"(&amap), "r"(j0), "r"(e0.y), "r"(e0.x * H), "r"(bar) : "memory");
          if (two) {
            asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                         ::"r"(st + PX_ZROW), "l"(z1 + (size_t)j0 * C), "r"(zb), "r"(bar), "l"(pol) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(st + 2 * PX_ZROW + 2048), "l"(&amap), "r"(j0), "r"(e1.y), "r"(e1.x * H), "r"(bar) : "memory");
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = idesc_tf32(128, 32);
    int g = 0, n = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++n) {
      const int buf = n & 1;
      mbar_wait(&acc_empty[buf], ((n >> 1) & 1) ^ 1);
      const uint32_t tb = tmem_base + PX_TM_ACC + buf * PX_ACC_COLS;
      for (int kb = 0; kb < nkb; ++kb, ++g) {
        const int s = g % PX_NST;
        mbar_wait(&ready[s], (g / PX_NST) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_hi = smem_u32(smem + s * PX_STAGE + 2 * PX_ZROW), b_lo = b_hi + PX_AB;
          const uint32_t ta = tmem_base + PX_TM_A + (g % PX_NSLOT) * 64;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t ah = ta + k * 8, al = ah + 32;
            const uint64_t dbh = smem_desc_sw128(b_hi + k * 32), dbl = smem_desc_sw128(b_lo + k * 32);
            px_mma_ts(tb + (kb / gsz) * 32, ah, dbh, idesc, (kb % gsz == 0 && k == 0) ? 0u : 1u);
            px_mma_ts(tb + 64, ah, dbl, idesc, (kb == 0 && k == 0) ? 0u : 1u);
            px_mma_ts(tb + 64, al, dbh, idesc, 1u);
          }
          mma_commit(&empty[s]);
          mma_commit(&ta_free[g % PX_NSLOT]);
          if (kb == nkb - 1) mma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else if (warp < PX_TW0) {
    // ---- epilogue warps; first the masked query rows that are needed: alpha row = 0 (ga.py:25) -> zero pair aggregate
    const int te = (warp - 2) * 32 + lane;
    for (int n = blockIdx.x; n < ndead; n += gridDim.x) {
      const int4 e = a.list[a.nrows - 1 - n];
      float* feat_row = a.feat + (size_t)e.z * a.feat_ld;
      float* alpha_row0 = a.alpha + ((size_t)(e.x * H) * L + e.y) * a.Lp;
      for (int o = te; o < H * C; o += 128) feat_row[o] = 0.f;
      for (int h = 0; h < H; ++h)
        for (int j = te; j < a.Lp; j += 128) alpha_row0[(size_t)h * L * a.Lp + j] = 0.f;
    }
    const int q = warp & 3, r = q >> 1, c = (q & 1) * 32 + lane;
    const int ngrp = (nkb + gsz - 1) / gsz;
    int n = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++n) {
      const int buf = n & 1;
      mbar_wait(&acc_full[buf], (n >> 1) & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + PX_TM_ACC + buf * PX_ACC_COLS + r * 16;
      float v[16], w[16];
      tmem_ld_32x16(trow, v);
      if (ngrp > 1) {
        tmem_ld_32x16(trow + 32, w);
#pragma unroll
        for (int h = 0; h < H; ++h) v[h] += w[h];
      }
      tmem_ld_32x16(trow + 64, w);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      const int idx = 2 * PX_TILE(tile) + r;
      if (idx < nlive) {
        float* feat_row = a.feat + (size_t)a.list[idx].z * a.feat_ld + c;
#pragma unroll
        for (int h = 0; h < H; ++h) feat_row[h * C] = v[h] + w[h];
      }
    }
    tc_fence_before();
  } else {
    // ---- transposers: thread = TMEM lane = (row of the tile, channel)
    const int gi = (warp - PX_TW0) >> 2, q = warp & 3, r = q >> 1, c = (q & 1) * 32 + lane;
    const int tg = ((warp - PX_TW0) & 3) * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + PX_TM_A;
    const int total = ((ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * nkb;      // chunks of this CTA
    for (int g = gi; g < total; g += PX_NG) {
      const int s = g % PX_NST, slot = g % PX_NSLOT;
      unsigned char* st = smem + s * PX_STAGE;
      mbar_wait(&full[s], (g / PX_NST) & 1);
      const float* zs = reinterpret_cast<const float*>(st + r * PX_ZROW) + c;
      float raw[32], lo[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) raw[k] = zs[k * C];
      {
        const float4* asrc = reinterpret_cast<const float4*>(st + 2 * PX_ZROW);
        float4* adst = reinterpret_cast<float4*>(st + 2 * PX_ZROW + PX_AB);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          const float4 v = asrc[tg + 128 * m];
          adst[tg + 128 * m] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
        }
      }
#pragma unroll
      for (int k = 0; k < 32; ++k) lo[k] = tf32_lo(raw[k]);
      mbar_wait(&ta_free[slot], ((g / PX_NSLOT) & 1) ^ 1);      // the MMAs of chunk g - 4 have read this TMEM slot
      tc_fence_after();
      px_st32(trow + slot * 64, raw);
      px_st32(trow + slot * 64 + 32, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ready[s]);
    }
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------ host side
static int g_sm_count = 0;

cudaError_t pair_stream_init() {
  cudaError_t e;
  int dev = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(pair_bias_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PB_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(pair_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PX_SMEM)) != cudaSuccess) return e;
  return cudaSuccess;
}

// bias[b][h][i][Lp] for complexes [b0, b0 + nb); z viewed as a 2-D fp32 matrix [(N*L*L) rows][64]
bool launch_pair_bias(int nb, int b0, int N, int L, int Lp, const float* z, const PairBiasPacked& pb, float* bias, cudaStream_t st) {
  CUtensorMap zmap;
  const uint64_t rows = (uint64_t)N * L * L;
  const uint32_t box_rows = rows < (uint64_t)PB_CJ ? (uint32_t)rows : (uint32_t)PB_CJ;
  if (!make_tmap(&zmap, z, rows, C, C, box_rows)) return false;
  ProfScope prof__(KK_OTHER, st);
  PairBiasArgs a{};
  a.box_bytes = (int)box_rows * 128;
  a.L = L; a.Lp = Lp; a.b0 = b0; a.nrows = nb * L; a.nchunk = (Lp + PB_CJ - 1) / PB_CJ; a.bias = bias;
  int grid = g_sm_count > 0 ? g_sm_count : 148;
  const int need = (a.nrows + PB_WARPS - 1) / PB_WARPS;
  if (grid > need) grid = need;
  pair_bias_kernel<<<grid, PB_THREADS, PB_SMEM, st>>>(zmap, pb, a);
  return true;
}

void launch_pair_rows_build(int nb, int b0, int L, const uint8_t* mask, const int* cidx, const PairRows& pr, cudaStream_t st,
                            bool compact) {
  ProfScope prof__(KK_OTHER, st);
  pair_rows_build_kernel<<<1, 1024, 0, st>>>(nb * L, L, b0, mask, cidx, compact ? 1 : 0, pr.list, pr.count);
}

// ------------------------------------------------------------------------------------------ context cache (first block)
// Inside the sampling loop the input of the FIRST GABlock changes only on the generated residues: x_0 = mixer(res_feat, s_t) and
// the frames (R_t, p_t) of a context residue are the same at every step (dpm_full.py:86-89, transition.py:99,158,176 keep the
// context).  For a context query i the logits against context keys are therefore loop invariants, and so is
//     P_i[h][c] = sum_{j context} a~_ij z_ijc ,   a~ = softmax over the context keys alone, with row maximum m~_i and sum S~_i
// (computed once per run: the logits kernel with the generated keys excluded, then pair_stream_kernel into a 768-wide cache).
// At a step with row maximum m_i and sum S_i over ALL keys, alpha_ij = a~_ij S~_i exp(m~_i - m_i) / S_i for context keys j, so
//     sum_j alpha_ij z_ij = P_i * [S~_i exp(m~_i - m_i) / S_i]  +  sum_{g generated} alpha_ig z_ig
// -- 16 instead of 256 rows of z per context query in C2 (m_i >= m~_i, so the factor never overflows).  ctx_delta_kernel does
// that for every context query row; the generated query rows go through pair_stream_kernel as usual (row list of the generated
// rows).  Same numbers up to fp32 reassociation; ABOPT_NO_CTXCACHE=1 streams all of z in the first block as well.
struct CtxDeltaArgs {
  int N, L, Lp;
  const float* z; const uint8_t* mask;
  float* alpha;                   // [N][H][L][Lp] of this step; rows of masked context queries are zeroed here (ga.py:25)
  const float* cache;             // [N * L][768]  P_i
  const float2* stats_ctx;        // [N][H][L]  (m~, S~)
  const float2* stats;            // [N][H][L]  (m, S) of this step
  const int* cidx;                // [N * L]  >= 0: generated row (handled by pair_stream_kernel)
  const int* rows;                // generated rows, complex by complex
  const int* first;               // [N] index of the complex's first generated row in `rows`
  const int* count;               // count[0] = number of generated rows
  float* feat;                    // [N * L][1824]
};

__global__ void __launch_bounds__(256, 3)
ctx_delta_kernel(const CtxDeltaArgs a) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
  const int L = a.L, Lp = a.Lp;
  const float l2e = 1.4426950408889634f;
  for (int r = warp; r < a.N * L; r += nwarp) {
    if (a.cidx[r] >= 0) continue;
    const int b = r / L, i = r - b * L;
    float* feat_row = a.feat + (size_t)r * NFEAT;
    float* alpha_row0 = a.alpha + ((size_t)(b * H) * L + i) * Lp;
    if (a.mask[r] == 0) {
      for (int o = lane; o < H * C; o += 32) feat_row[o] = 0.f;
      for (int h = 0; h < H; ++h)
        for (int j = lane; j < Lp; j += 32) alpha_row0[(size_t)h * L * Lp + j] = 0.f;
      continue;
    }
    // per head: weight of the cached context part
    float wl = 0.f;
    if (lane < H) {
      const size_t si = (size_t)(b * H + lane) * L + i;
      const float2 sc = a.stats_ctx[si], st = a.stats[si];
      float e;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((sc.x - st.x) * l2e));
      wl = sc.y * e / st.y;
    }
    float2 acc[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float w = __shfl_sync(0xffffffffu, wl, h);
      const float2 p = *reinterpret_cast<const float2*>(a.cache + (size_t)r * (H * C) + h * C + 2 * lane);
      acc[h] = make_float2(p.x * w, p.y * w);
    }
    const int k0 = a.first[b], k1 = (b + 1 < a.N) ? a.first[b + 1] : a.count[0];
    for (int kb = k0; kb < k1; kb += 32) {
      const int nk = (k1 - kb < 32) ? (k1 - kb) : 32;
      const int g = (lane < nk) ? a.rows[kb + lane] - b * L : 0;       // lane l holds the l-th generated key of this pass
      float al[H];
#pragma unroll
      for (int h = 0; h < H; ++h) al[h] = (lane < nk) ? alpha_row0[(size_t)h * L * Lp + g] : 0.f;
      // eight z rows in flight per warp (the kernel is latency-, not bandwidth-bound: ~11 KB per query row)
      for (int k8 = 0; k8 < nk; k8 += 8) {
        float2 zv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int gk = __shfl_sync(0xffffffffu, g, (k8 + u) & 31);
          // (beyond nk: lane's g is 0, a valid row, and its weight below is 0 -- an unconditional load keeps all eight in flight)
          zv[u] = __ldg(reinterpret_cast<const float2*>(a.z + (((size_t)b * L + i) * L + gk) * C + 2 * lane));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
          for (int h = 0; h < H; ++h) {
            const float w = __shfl_sync(0xffffffffu, al[h], (k8 + u) & 31);      // (0 beyond nk)
            acc[h].x = fmaf(w, zv[u].x, acc[h].x);
            acc[h].y = fmaf(w, zv[u].y, acc[h].y);
          }
        }
      }
    }
#pragma unroll
    for (int h = 0; h < H; ++h) *reinterpret_cast<float2*>(feat_row + h * C + 2 * lane) = acc[h];
  }
}

void launch_ctx_delta(int N, int L, int Lp, const float* z, const uint8_t* mask, float* alpha, const float* cache,
                      const float2* stats_ctx, const float2* stats, const int* cidx, const int* rows, const int* first,
                      const int* count, float* feat, cudaStream_t st) {
  ProfScope prof__(KK_CTX, st);
  const CtxDeltaArgs a{N, L, Lp, z, mask, alpha, cache, stats_ctx, stats, cidx, rows, first, count, feat};
  int grid = (N * L + 7) / 8;
  const int cap = (g_sm_count > 0 ? g_sm_count : 148) * 3;      // one resident wave (3 CTAs per SM), rows strided over the warps
  if (grid > cap) grid = cap;
  ctx_delta_kernel<<<grid, 256, 0, st>>>(a);
}

bool launch_pair_stream(int nb, int b0, int L, int Lp, const float* z, float* alpha, float* feat, const PairRows& pr, cudaStream_t st,
                        int feat_ld, bool partial) {
  CUtensorMap amap;
  // alpha as a 3-D tensor [nb * H][L queries][Lp keys], 128-byte swizzle; box = [12 heads][1 query][32 keys] = 12 rows of the
  // K-major B operand (out-of-range keys arrive as zeros)
  if (!make_tmap_3d_sw128(&amap, alpha, (uint64_t)Lp, (uint64_t)L, (uint64_t)nb * H, PX_CJ, 1, H)) return false;
  ProfScope prof__(partial ? KK_PAIR_PART : KK_PAIR, st);
  PairRowsArgs a{};
  a.L = L; a.Lp = Lp; a.b0 = b0; a.nrows = nb * L; a.nchunk = (L + PX_CJ - 1) / PX_CJ;
  a.z = z; a.alpha = alpha; a.feat = feat; a.feat_ld = feat_ld; a.list = pr.list; a.count = pr.count;
  int grid = g_sm_count > 0 ? g_sm_count : 148;
  const int need = (a.nrows + 1) / 2;
  if (grid > need) grid = need;
  pair_stream_kernel<<<grid, PX_THREADS, PX_SMEM, st>>>(amap, a);
  return true;
}

}  // namespace abopt
