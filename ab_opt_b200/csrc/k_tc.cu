// tcgen05 "3xTF32" GEMM:  D[M][N] = A[M][K] * B[N][K]^T  with fp32-grade accuracy on the 5th-gen tensor cores.
//
// Both operands arrive pre-split into tf32-exact parts, x = hi + lo (tc::split_tf32), and every 128x BN x 32 k-block
// issues the three products  A_hi*B_hi + A_hi*B_lo + A_lo*B_hi  into one TMEM accumulator (the dropped lo*lo term is
// ~2^-22 relative).  The reference runs these linears in true fp32 (allow_tf32 is off in its inference runners), and
// plain TF32 (10-bit mantissa) would break the 1e-4 parity bar, hence the split.
//
// Structure (one CTA per 128 x BN output tile, 192 threads):
//   warp 0      : TMA producer  -- cp.async.bulk.tensor, 128B-swizzled [rows][32 tf32] boxes, mbarrier expect_tx
//   warp 1      : TMEM allocator + MMA issuer (one elected lane issues tcgen05.mma.kind::tf32, commits to mbarriers)
//   warps 2..5  : epilogue      -- tcgen05.ld (thread = output row), fused epilogue functor, global stores
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include "tc.cuh"
#include "params.cuh"
#include "kernels.h"

namespace abopt {

using namespace tc;

constexpr int G_BM = 128, G_BK = 32, G_THREADS = 192;

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = G_BM * G_BK * 4;       // 16 KB
  static constexpr int B_BYTES = BN * G_BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;    // barriers (<= 11 x 8 B + slot) + slack for 1024 B alignment
  static_assert(BN % 32 == 0, "epilogue reads TMEM in 32-column chunks");    // barriers + slack for 1024 B alignment
};

// ---- epilogues: called once per output row with the BN accumulators of that row in registers
struct EpiPlain {            // D = acc (+ bias)
  float* D; int ldd; const float* bias;
  template <int BN>
  __device__ __forceinline__ void operator()(int row, int n0, int N, float (&v)[BN]) const {
    float* dst = D + (size_t)row * ldd + n0;
#pragma unroll
    for (int c = 0; c < BN; c += 4) {
      if (n0 + c < N) {
        float4 o = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        if (bias) { o.x += bias[n0 + c]; o.y += bias[n0 + c + 1]; o.z += bias[n0 + c + 2]; o.w += bias[n0 + c + 3]; }
        *reinterpret_cast<float4*>(dst + c) = o;
      }
    }
  }
};

// GABlock projections, packed for the tensor-core attention kernels (k_attn_tc.cu).  Per (complex b, head h, residue r):
//   QA[b][h][r][64] = [ q / sqrt(32) (32) | global query points (24) | 0 (8) ]                    ga.py:82-85,96-99
//   KB[b][h][r][64] = [ k (32)            | -2 c_h * global key points (24) | 0 (8) ]             ga.py:83,102-105
//   rq[b][h][r] = c_h |query points|^2,  rk[b][h][r] = c_h |key points|^2,   c_h = -softplus(coef_h) sqrt(2/(9*8)) / 2
// so that  QA . KB + rq + rk = node logits + spatial logits  (|q - k|^2 expanded; ga.py:108-111).  The tf32 "lo"
// planes of all of them are built on chip by their consumers.  Values go out TRANSPOSED (key index contiguous), the K-major B
// operand of aggr_persist_kernel:
//   VT[b][h][n][r] = value channel n (n < 32) | global value point coordinate n - 32 (32 <= n < 56); rows 56..63 stay 0.
struct EpiProjPack {
  const float* R; const float* t; const float* coef;
  float* QA; float* QA_lo; float* KB; float* KB_lo; float* rq; float* rk; float* VT;
  int L, Lp;
  int qk_lo;            // 1: also write QA_lo / KB_lo (only the non-persistent logits kernels read them)
  const int* rows;      // optional: the A operand holds a compact list of residue rows (row k of it = residue row rows[k],
  const int* count;     //           count[0] of them, device side); the packed outputs go to the residues' own places
  int nsplit;           // work unit = (row tile, 1 / nsplit of the 21 column tiles): 1 for a full batch (x stays resident over all
                        // column tiles), 7 for a short row list, where one CTA per row tile would leave most SMs idle
};

// Accuracy note (measured on B200, scripts/debug_gemm.py): the tensor core TRUNCATES the fp32 accumulator on every
// tcgen05.mma, a systematic -2^-24 relative bias per accumulation that grows linearly with K (4e-5 at K = 1824).
// Two counter-measures keep the result fp32-grade:
//   (1) the large product A_hi*B_hi and the small corrections A_hi*B_lo + A_lo*B_hi go to SEPARATE TMEM accumulators
//       (the corrections are 2^-11 smaller, so their truncation is irrelevant), summed in fp32 registers at the end;
//   (2) "promotion": every KCH k-blocks the accumulators are drained into fp32 registers (round-to-nearest adds on the
//       CUDA cores) while the MMAs continue into the other half of a double-buffered TMEM allocation.
template <int BN, int STAGES, int KCH, class Epi>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm3x_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
              const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, int M, int N, int K, Epi epi) {
  using S = GemmSmem<BN, STAGES>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;        // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * G_BM;
  const int nkb = K / G_BK;
  const int nchunk = (nkb + KCH - 1) / KCH;
  constexpr uint32_t ACC_COLS = 2 * BN;                                  // main | small
  constexpr uint32_t TMEM_COLS = (2 * ACC_COLS <= 256) ? 256 : 512;      // double buffered

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); }
    mbar_fence_init();
    tma_prefetch_desc(&tmAh); tma_prefetch_desc(&tmAl); tma_prefetch_desc(&tmBh); tma_prefetch_desc(&tmBl);
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);                                  // slot free (first round passes immediately)
        unsigned char* st = smem + s * S::STAGE_BYTES;
        mbar_expect_tx(&full[s], S::STAGE_BYTES);
        tma_load_2d(st, &tmAh, kb * G_BK, m0, &full[s]);
        tma_load_2d(st + S::A_BYTES, &tmAl, kb * G_BK, m0, &full[s]);
        tma_load_2d(st + 2 * S::A_BYTES, &tmBh, kb * G_BK, n0, &full[s]);
        tma_load_2d(st + 2 * S::A_BYTES + S::B_BYTES, &tmBl, kb * G_BK, n0, &full[s]);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = idesc_tf32(G_BM, BN);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      const int c = kb / KCH, buf = c & 1;
      const bool first = (kb % KCH) == 0, last = (kb % KCH) == KCH - 1 || kb == nkb - 1;
      if (first && c >= 2) { mbar_wait(&tmem_empty[buf], ((c >> 1) - 1) & 1); }     // epilogue drained this buffer
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_hi = smem_u32(smem + s * S::STAGE_BYTES), a_lo = a_hi + S::A_BYTES;
        const uint32_t b_hi = a_hi + 2 * S::A_BYTES, b_lo = b_hi + S::B_BYTES;
        const uint32_t d_main = tmem_base + buf * ACC_COLS, d_small = d_main + BN;
#pragma unroll
        for (int k = 0; k < G_BK / 8; ++k) {                           // UMMA_K = 8 tf32 = 32 B inside the 128 B swizzle row
          const uint64_t dah = smem_desc_sw128(a_hi + k * 32), dal = smem_desc_sw128(a_lo + k * 32);
          const uint64_t dbh = smem_desc_sw128(b_hi + k * 32), dbl = smem_desc_sw128(b_lo + k * 32);
          const uint32_t acc = (first && k == 0) ? 0u : 1u;
          mma_tf32(d_main, dah, dbh, idesc, acc);
          mma_tf32(d_small, dah, dbl, idesc, acc);
          mma_tf32(d_small, dal, dbh, idesc, 1u);
        }
        mma_commit(&empty[s]);                                         // smem slot reusable once these MMAs retire
        if (last) mma_commit(&tmem_full[buf]);                         // this chunk's accumulators are complete
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                                            // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
    float v[BN];
#pragma unroll
    for (int i = 0; i < BN; ++i) v[i] = 0.f;
    for (int c = 0; c < nchunk; ++c) {
      const int buf = c & 1;
      mbar_wait(&tmem_full[buf], (c >> 1) & 1);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + buf * ACC_COLS;
#pragma unroll
      for (int cc = 0; cc < BN; cc += 32) {
        float tm[32], ts[32];
        tmem_ld_32x32(tbase + cc, tm);
        tmem_ld_32x32(tbase + BN + cc, ts);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[cc + i] += tm[i] + ts[i];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
    if (row < M) epi.template operator()<BN>(row, n0, N, v);
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ---------------------------------------------------------------- persistent projection GEMM
// proj_ts_kernel: the six GABlock projections, (B*L, 128) x (128, 2016), with the EpsProjPack packing.  What bounded its
// predecessor (round 2 timeline: one k-block of 12 SS-mode tcgen05.mma every ~1200 cycles, ~100 cycles per 128x96x8 instruction
// against the 48 of the issue formula) was the shared-memory operand traffic of the MMAs: A and B were both read from shared
// memory, 7 KB per instruction.  Here the A operand lives in TENSOR MEMORY:
//   * x of the CTA's 128-row tile is loaded ONCE from global memory into registers by the first epilogue group (thread = row),
//     split into its tf32 hi / lo planes on the fly and stored to TMEM columns [0,128) | [128,256) (tcgen05.st); no x in shared
//     memory, no x_lo in global memory;
//   * the MMAs are issued in TS mode (A from TMEM, B from shared memory): 2-3 KB of shared-memory reads per instruction;
//   * shared memory is all weight ring: 8 stages of [64 columns][32 k] hi | lo (16 KB), streamed from L2 by TMA;
//   * column tiles of 64 (q / k / v: two heads) and 48 (points: two heads x 8 points x 3), 36 per row tile, two 128-column
//     TMEM accumulator sets (main | corrections) filled alternately, two epilogue groups packing them.
// warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 / 6-9 = epilogue groups 0 / 1 (group 0 also loads A).
constexpr int PP_THREADS = 320, PP_KB = F / G_BK, PP_ST = 8;
constexpr int PP_NT64 = 3 * H * D / 64, PP_NT48 = 3 * H * P * 3 / 48, PP_NT = PP_NT64 + PP_NT48;      // 18 + 18 column tiles
constexpr int PP_B_BYTES = 64 * G_BK * 4;                   // 8 KB: 64 weight rows x 32 tf32 (the 48-wide tiles use the first 48)
constexpr int PP_STAGE = 2 * PP_B_BYTES;                    // 16 KB: weights hi | lo of one k-block
constexpr int PP_BAR_OFF = PP_ST * PP_STAGE;
constexpr int PP_SMEM = PP_BAR_OFF + 256 + 1024;
constexpr uint32_t PP_TM_A = 0, PP_TM_ACC = 256;            // TMEM columns: x hi (128) | x lo (128) | 2 x (main 64 | corrections 64)
static_assert(PP_NT64 * 64 == 3 * H * D && PP_NT48 * 48 == 3 * H * P * 3 && PP_KB * G_BK == F, "projection tiling");

__device__ __forceinline__ int pp_tile_n0(int nt) { return nt < PP_NT64 ? nt * 64 : OFF_QP + (nt - PP_NT64) * 48; }

// D[tmem] (+)= A[tmem] * B[smem]; one thread issues
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x8_nw(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// W columns (a multiple of 8) of the main and the correction accumulator, summed
template <int W>
__device__ __forceinline__ void tmem_ld_sum(uint32_t t_main, uint32_t t_small, float (&v)[W]) {
  float m[W], c[W];
#pragma unroll
  for (int i = 0; i < W; i += 8) { tmem_ld_32x8_nw(t_main + i, m + i); tmem_ld_32x8_nw(t_small + i, c + i); }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < W; ++i) v[i] = m[i] + c[i];
}
__device__ __forceinline__ void st_v8(float* p, const float (&w)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(w[0]), "f"(w[1]), "f"(w[2]), "f"(w[3]),
               "f"(w[4]), "f"(w[5]), "f"(w[6]), "f"(w[7]) : "memory");
}
__device__ __forceinline__ void st_v8_hi_lo(float* hi, float* lo, const float (&w)[8], bool with_lo = true) {
  st_v8(hi, w);
  if (!with_lo) return;
  float l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) l[i] = tf32_lo(w[i]);
  st_v8(lo, l);
}

__global__ void __launch_bounds__(PP_THREADS, 1)
proj_ts_kernel(const float* __restrict__ xg, const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
               int M, EpiProjPack ep) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + PP_BAR_OFF);
  uint64_t* a_empty = a_full + 1;
  uint64_t* b_full = a_empty + 1;            // [PP_ST]
  uint64_t* b_empty = b_full + PP_ST;        // [PP_ST]
  uint64_t* tmem_full = b_empty + PP_ST;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Mrows = ep.rows ? min(M, ep.count[0]) : M;      // (compact list: the row count lives on the device)
  const int nmt = (Mrows + G_BM - 1) / G_BM;
  const int nunits = nmt * ep.nsplit;

  if (threadIdx.x == 0) {
    mbar_init(a_full, 4); mbar_init(a_empty, 1);
    for (int s = 0; s < PP_ST; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); }
    mbar_fence_init();
    tma_prefetch_desc(&tmBh); tma_prefetch_desc(&tmBl);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: the weight stream =====================
    if (elect_one()) {
      int g = 0;
      for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
        const int part = u % ep.nsplit;
        const int nt0 = part * PP_NT / ep.nsplit, nt1 = (part + 1) * PP_NT / ep.nsplit;
        for (int nt = nt0; nt < nt1; ++nt) {
          const int n0 = pp_tile_n0(nt);
          for (int kb = 0; kb < PP_KB; ++kb, ++g) {
            const int s = g % PP_ST;
            mbar_wait(&b_empty[s], ((g / PP_ST) & 1) ^ 1);
            unsigned char* st = smem + s * PP_STAGE;
            mbar_expect_tx(&b_full[s], PP_STAGE);                       // (rows past the end of W arrive as zeros, and count)
            tma_load_2d(st, &tmBh, kb * G_BK, n0, &b_full[s]);
            tma_load_2d(st + PP_B_BYTES, &tmBl, kb * G_BK, n0, &b_full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (A from tensor memory) =====================
    constexpr uint32_t idesc64 = idesc_tf32(G_BM, 64), idesc48 = idesc_tf32(G_BM, 48);
    int na = 0, g = 0, n = 0;
    for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++na) {
      const int part = u % ep.nsplit;
      const int nt0 = part * PP_NT / ep.nsplit, nt1 = (part + 1) * PP_NT / ep.nsplit;
      mbar_wait(a_full, na & 1);                                        // x hi | lo of this row tile are in TMEM
      tc_fence_after();
      for (int nt = nt0; nt < nt1; ++nt, ++n) {
        const int buf = n & 1;
        const uint32_t idesc = nt < PP_NT64 ? idesc64 : idesc48;
        mbar_wait(&tmem_empty[buf], ((n >> 1) & 1) ^ 1);              // epilogue group `buf` drained tile n - 2
        for (int kb = 0; kb < PP_KB; ++kb, ++g) {
          const int s = g % PP_ST;
          mbar_wait(&b_full[s], (g / PP_ST) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t b_hi = smem_u32(smem + s * PP_STAGE);      // W_lo box at + PP_B_BYTES
            const uint32_t d_main = tmem_base + PP_TM_ACC + buf * 128, d_small = d_main + 64;
#pragma unroll
            for (int k = 0; k < G_BK / 8; ++k) {
              const uint32_t ah = tmem_base + PP_TM_A + kb * G_BK + k * 8, al = ah + F;
              const uint64_t dbh = smem_desc_sw128(b_hi + k * 32);
              const uint32_t acc = (kb == 0 && k == 0) ? 0u : 1u;
              // x_hi . [W_hi | W_lo] -> [main | corrections] as ONE 128-wide instruction: the lo box follows the hi box in the stage
              // (64 + 64 rows) and the corrections accumulator follows the main one in TMEM.  A 128 x 64 x 8 instruction costs ~45 clk,
              // a 128 x 128 x 8 one ~77, and this kernel is bound by MMA issue: 60.0 -> 56.3 us per launch.  (48-wide tiles compute 16
              // columns nobody reads.)
              mma_tf32_ts(d_main, ah, dbh, idesc_tf32(G_BM, 128), acc);
              mma_tf32_ts(d_small, al, dbh, idesc, 1u);
            }
            mma_commit(&b_empty[s]);
            if (kb == PP_KB - 1) {
              mma_commit(&tmem_full[buf]);
              if (nt == nt1 - 1) mma_commit(a_empty);
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue groups =====================
    const int q = warp & 3;                                            // TMEM lane quarter this warp may access
    const int gp = (warp - 2) >> 2;                                    // group: takes the tiles with n % 2 == gp
    const int L = ep.L, Lp = ep.Lp;
    int n = 0, na = 0;
    for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++na) {
      const int mt = u / ep.nsplit, part = u - mt * ep.nsplit;
      const int nt0 = part * PP_NT / ep.nsplit, nt1 = (part + 1) * PP_NT / ep.nsplit;
      const int row = mt * G_BM + q * 32 + lane;
      const bool valid = row < Mrows;
      if (gp == 0) {
        // ---- A operand: this thread's row of x -> tf32 hi | lo planes in TMEM (lane = row, column = k)
        mbar_wait(a_empty, (na & 1) ^ 1);                               // the MMAs of the previous unit have read it
        tc_fence_after();
        const float4* xr = reinterpret_cast<const float4*>(xg + (size_t)(valid ? row : 0) * F);
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + PP_TM_A;
#pragma unroll 1
        for (int c = 0; c < F; c += 32) {
          float hi[32], lo[32];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float4 v = valid ? __ldg(xr + (c >> 2) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
            hi[4 * e] = v.x; hi[4 * e + 1] = v.y; hi[4 * e + 2] = v.z; hi[4 * e + 3] = v.w;
          }
#pragma unroll
          for (int e = 0; e < 32; ++e) lo[e] = tf32_lo(hi[e]);
          tmem_st_32x32(ta + c, hi);                                    // (raw fp32: the tensor core ignores the low 13 mantissa bits)
          tmem_st_32x32(ta + F + c, lo);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full);
      }
      const int rr = valid ? (ep.rows ? ep.rows[row] : row) : 0;
      const int b = rr / L, r = rr - b * L;
      float Rm[9], tv[3];
#pragma unroll
      for (int i = 0; i < 9; ++i) Rm[i] = __ldg(ep.R + (size_t)rr * 9 + i);
#pragma unroll
      for (int i = 0; i < 3; ++i) tv[i] = __ldg(ep.t + (size_t)rr * 3 + i);
      for (int nt = nt0; nt < nt1; ++nt, ++n) {
        if ((n & 1) != gp) continue;
        mbar_wait(&tmem_full[gp], (n >> 1) & 1);
        tc_fence_after();
        const uint32_t tm = tmem_base + ((uint32_t)(q * 32) << 16) + PP_TM_ACC + gp * 128, ts = tm + 64;
        const int n0 = pp_tile_n0(nt);
        if (n0 < OFF_V) {
          // ---- q or k channels: 2 heads x 32                                      ga.py:82-85
          const bool is_q = n0 < OFF_K;
          const int h0 = (is_q ? n0 : n0 - OFF_K) / D;
          const float sc = is_q ? 0.17677669529663687f : 1.f;          // 1 / sqrt(32) folded into q
          float* dst = is_q ? ep.QA : ep.KB;
          float* dlo = is_q ? ep.QA_lo : ep.KB_lo;
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {
            float v[32];
            tmem_ld_sum<32>(tm + hh * D, ts + hh * D, v);
            if (valid) {
              const size_t o = ((size_t)(b * H + h0 + hh) * L + r) * 64;
#pragma unroll
              for (int c = 0; c < D; c += 8) {
                float w[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) w[e] = v[c + e] * sc;
                st_v8_hi_lo(dst + o + c, dlo + o + c, w, ep.qk_lo != 0);
              }
            }
          }
        } else if (n0 < OFF_QP) {
          // ---- value channels: 2 heads x 32, stored transposed (key index contiguous)  ga.py:122
          const int h0 = (n0 - OFF_V) / D;
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {
            float v[32];
            tmem_ld_sum<32>(tm + hh * D, ts + hh * D, v);
            if (valid) {
              const size_t o = ((size_t)(b * H + h0 + hh) * 64) * Lp + r;
#pragma unroll
              for (int c = 0; c < D; ++c) ep.VT[o + (size_t)c * Lp] = v[c];
            }
          }
        } else {
          // ---- points: 2 heads x 8 points x 3, local -> global q = R p + t      geometry.py:72-91
          const int kind = n0 < OFF_KP ? 0 : (n0 < OFF_VP ? 1 : 2);      // query / key / value points
          const int h0 = (n0 - (kind == 0 ? OFF_QP : (kind == 1 ? OFF_KP : OFF_VP))) / (P * 3);
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {
            float v[24];
            tmem_ld_sum<24>(tm + hh * P * 3, ts + hh * P * 3, v);
            if (valid) {
#pragma unroll
              for (int p = 0; p < P * 3; p += 3) {
                const float x = v[p], y = v[p + 1], z = v[p + 2];
                v[p + 0] = Rm[0] * x + Rm[1] * y + Rm[2] * z + tv[0];
                v[p + 1] = Rm[3] * x + Rm[4] * y + Rm[5] * z + tv[1];
                v[p + 2] = Rm[6] * x + Rm[7] * y + Rm[8] * z + tv[2];
              }
              const int h = h0 + hh;
              if (kind < 2) {
                // query / key points -> QA / KB columns 32..63 + norm terms           ga.py:96-111
                float* dst = kind == 0 ? ep.QA : ep.KB;
                float* dlo = kind == 0 ? ep.QA_lo : ep.KB_lo;
                float* rn = kind == 0 ? ep.rq : ep.rk;
                const float ch = __ldg(ep.coef + h);
                const float sc = kind == 0 ? 1.f : -2.f * ch;
                float n2 = 0.f;
#pragma unroll
                for (int c = 0; c < P * 3; ++c) n2 = fmaf(v[c], v[c], n2);
                rn[(size_t)(b * H + h) * L + r] = ch * n2;
                const size_t o = ((size_t)(b * H + h) * L + r) * 64 + D;
#pragma unroll
                for (int c = 0; c < 32; c += 8) {
                  float w[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) w[e] = (c + e < P * 3) ? v[c + e] * sc : 0.f;
                  st_v8_hi_lo(dst + o + c, dlo + o + c, w, ep.qk_lo != 0);
                }
              } else {
                const size_t o = ((size_t)(b * H + h) * 64 + D) * Lp + r;
#pragma unroll
                for (int c = 0; c < P * 3; ++c) ep.VT[o + (size_t)c * Lp] = v[c];
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[gp]);
      }
    }
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

cudaError_t tc_init() {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return e;
    if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
    g_encode = (PFN_encodeTiled)fn;
  }
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(gemm3x_kernel<128, 3, 8, EpiPlain>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<128, 3>::TOTAL)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(proj_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PP_SMEM)) != cudaSuccess) return e;
  return cudaSuccess;
}

// 2-D fp32 row-major [rows][cols] with row pitch ld (floats); box = [box_rows][32 floats], 128 B swizzle.
// Descriptors are cached per (address, shape): the workspace buffers and weights they describe are long-lived, and
// encoding costs a few microseconds of host time per call.
struct TmapKey {
  const void* base; uint64_t rows, cols, ld; uint32_t box_rows;
  bool operator<(const TmapKey& o) const {
    return std::tie(base, rows, cols, ld, box_rows) < std::tie(o.base, o.rows, o.cols, o.ld, o.box_rows);
  }
};
static std::map<TmapKey, CUtensorMap> g_tmaps;
static std::mutex g_tmaps_mu;
bool make_tmap(CUtensorMap* m, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  std::lock_guard<std::mutex> lk(g_tmaps_mu);
  const TmapKey key{base, rows, cols, ld, box_rows};
  auto it = g_tmaps.find(key);
  if (it != g_tmaps.end()) { *m = it->second; return true; }
  if (!make_tmap_2d(m, base, rows, cols, ld, box_rows, G_BK)) return false;
  if (g_tmaps.size() > 4096) g_tmaps.clear();
  g_tmaps[key] = *m;
  return true;
}

bool make_tmap_2d(CUtensorMap* m, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  if (!g_encode) return false;
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// 3-D fp32 tensor [d2][d1][d0] (d0 contiguous, dense), box [1][box1][box0], 128 B swizzle (box0 * 4 <= 128); cached
bool make_tmap_3d(CUtensorMap* m, const float* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box0, uint32_t box1) {
  std::lock_guard<std::mutex> lk(g_tmaps_mu);
  const TmapKey key{base, d1 * 4096 + d2, d0, ((uint64_t)box0 << 32) | 0x3D, box1};
  auto it = g_tmaps.find(key);
  if (it != g_tmaps.end()) { *m = it->second; return true; }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * sizeof(float), d0 * d1 * sizeof(float)};
  cuuint32_t box[3] = {box0, d1 < box1 ? (cuuint32_t)d1 : box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if (!g_encode) return false;
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  if (g_tmaps.size() > 4096) g_tmaps.clear();
  g_tmaps[key] = *m;
  return true;
}

// plain (unswizzled) 3-D fp32 tensor [d2][d1][d0] (d0 contiguous, dense), box [box2][box1][box0], zero fill out of range; cached
bool make_tmap_3d_plain(CUtensorMap* m, const float* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box0, uint32_t box1, uint32_t box2) {
  std::lock_guard<std::mutex> lk(g_tmaps_mu);
  const TmapKey key{base, d1 * 4096 + d2, d0, ((uint64_t)box0 << 32) | ((uint64_t)box2 << 8) | 0x3E, box1};
  auto it = g_tmaps.find(key);
  if (it != g_tmaps.end()) { *m = it->second; return true; }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * sizeof(float), d0 * d1 * sizeof(float)};
  cuuint32_t box[3] = {box0, d1 < box1 ? (cuuint32_t)d1 : box1, d2 < box2 ? (cuuint32_t)d2 : box2};
  cuuint32_t estr[3] = {1, 1, 1};
  if (!g_encode) return false;
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  if (g_tmaps.size() > 4096) g_tmaps.clear();
  g_tmaps[key] = *m;
  return true;
}

// the same 3-D tensor with the 128-byte swizzle (box0 * 4 bytes <= 128): rows of box [box2][box1] land 128 B apart
bool make_tmap_3d_sw128(CUtensorMap* m, const float* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box0, uint32_t box1, uint32_t box2) {
  std::lock_guard<std::mutex> lk(g_tmaps_mu);
  const TmapKey key{base, d1 * 4096 + d2, d0, ((uint64_t)box0 << 32) | ((uint64_t)box2 << 8) | 0x3C, box1};
  auto it = g_tmaps.find(key);
  if (it != g_tmaps.end()) { *m = it->second; return true; }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * sizeof(float), d0 * d1 * sizeof(float)};
  cuuint32_t box[3] = {box0, d1 < box1 ? (cuuint32_t)d1 : box1, d2 < box2 ? (cuuint32_t)d2 : box2};
  cuuint32_t estr[3] = {1, 1, 1};
  if (!g_encode) return false;
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  if (g_tmaps.size() > 4096) g_tmaps.clear();
  g_tmaps[key] = *m;
  return true;
}

// plain (unswizzled) 2-D fp32 tensor map, used for L2 prefetches only
bool make_tmap_plain(CUtensorMap* m, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  if (!g_encode) return false;
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

__global__ void split_kernel(const float* __restrict__ in, float* __restrict__ hi, float* __restrict__ lo, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float h, l;
    split_tf32(in[i], h, l);
    hi[i] = h; lo[i] = l;
  }
}
// lo = x - trunc_tf32(x): the second operand plane of the 3xTF32 scheme.  The FIRST plane is the raw fp32 tensor itself:
// tcgen05.mma.kind::tf32 ignores the low 13 mantissa bits of its 32-bit inputs, i.e. it sees trunc_tf32(x).
__global__ void lo_kernel(const float4* __restrict__ in, float4* __restrict__ lo, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = in[i];
    float h, l0, l1, l2, l3;
    split_tf32(v.x, h, l0); split_tf32(v.y, h, l1); split_tf32(v.z, h, l2); split_tf32(v.w, h, l3);
    lo[i] = make_float4(l0, l1, l2, l3);
  }
}
void launch_lo(const float* in, float* lo, size_t n, cudaStream_t st) {      // n % 4 == 0
  ProfScope prof__(KK_OTHER, st);
  const size_t n4 = n / 4;
  int grid = (int)((n4 + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < 1) grid = 1;
  lo_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(lo), n4);
}
void launch_split(const float* in, float* hi, float* lo, size_t n, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  split_kernel<<<grid, 256, 0, st>>>(in, hi, lo, n);
}

// Th[c][r] = W[r][c] (raw fp32 = the tf32 "hi" plane), Tl = its lo plane: a weight as the K-major B operand of an input-gradient GEMM
__global__ void transpose_split_kernel(const float* __restrict__ W, int R, int Cc, float* __restrict__ Th, float* __restrict__ Tl) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= (size_t)R * Cc) return;
  const int r = (int)(idx / Cc), c = (int)(idx - (size_t)r * Cc);
  const float v = W[idx];
  Th[(size_t)c * R + r] = v;
  Tl[(size_t)c * R + r] = tf32_lo(v);
}
void launch_transpose_split(const float* W, int R, int Cc, float* Th, float* Tl, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  transpose_split_kernel<<<(unsigned)(((size_t)R * Cc + 255) / 256), 256, 0, st>>>(W, R, Cc, Th, Tl);
}

// D[M][N] = A * B^T (+bias), operands given as hi / lo planes.  K % 32 == 0, N % 4 == 0.
bool launch_gemm3x_plain(int M, int N, int K, const float* Ah, const float* Al, int lda, const float* Bh, const float* Bl, int ldb,
                         float* D, int ldd, const float* bias, cudaStream_t st) {
  CUtensorMap a_h, a_l, b_h, b_l;
  if (!make_tmap(&a_h, Ah, M, K, lda, G_BM) || !make_tmap(&a_l, Al, M, K, lda, G_BM) || !make_tmap(&b_h, Bh, N, K, ldb, 128) ||
      !make_tmap(&b_l, Bl, N, K, ldb, 128))
    return false;
  ProfScope prof__(KK_TAIL, st);
  dim3 grid((N + 127) / 128, (M + G_BM - 1) / G_BM);
  gemm3x_kernel<128, 3, 8, EpiPlain><<<grid, G_THREADS, GemmSmem<128, 3>::TOTAL, st>>>(a_h, a_l, b_h, b_l, M, N, K, EpiPlain{D, ldd, bias});
  return true;
}

// the six GABlock projections as one GEMM, outputs packed for the tensor-core attention kernels (see EpiProjPack)
bool launch_proj_pack(int M, int L, int Lp, const float* xh, const float* xl, const float* Wh, const float* Wl, const float* R, const float* t,
                      const float* coef, const AttnOperands& op, cudaStream_t st, const int* rows, const int* count) {
  CUtensorMap b_h, b_l;
  (void)xl;      // (the lo plane of x is built on chip)
  if (!make_tmap(&b_h, Wh, NPROJ, F, F, 64) || !make_tmap(&b_l, Wl, NPROJ, F, F, 64)) return false;
  ProfScope prof__(KK_PROJ, st);
  const EpiProjPack ep{R, t, coef, op.QA, op.QA_lo, op.KB, op.KB_lo, op.rq, op.rk, op.VT, L, Lp, attn_needs_qk_lo(L) ? 1 : 0, rows, count, rows ? 9 : 1};
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int nunits = ((M + G_BM - 1) / G_BM) * ep.nsplit;
  proj_ts_kernel<<<nunits < sms ? nunits : sms, PP_THREADS, PP_SMEM, st>>>(xh, b_h, b_l, M, ep);
  return true;
}

}  // namespace abopt
