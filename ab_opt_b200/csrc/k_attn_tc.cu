// Attention logits and the softmax over key residues on the 5th-gen tensor cores (attn_logits_persist_kernel), and the
// node / point aggregation GEMM (aggr_persist_kernel).
//
// For one (complex b, head h, tile of 128 query residues) the logits kernel computes
//   D[i][j] = QA[i] . KB[j]            64-wide contraction: q.k / sqrt(32)  and  -2 c_h qp_i . kp_j   (3xTF32, TMEM accumulator)
//   l[i][j] = ((D + pair_bias(i, j)) + (rq[i] + rk[j])) * sqrt(1/3) - 1e5 [key j masked]              ga.py:81-112,166,23
//   alpha[i][:] = softmax_j l[i][:]                                                                       ga.py:24
// QA / KB / rq / rk come packed from the projection GEMM epilogue (k_tc.cu: EpiProjPack); pair_bias is the hoisted
// z . W_b (k_pair.cu: pair_bias_kernel).
#include <cstdlib>
#include "tc.cuh"
#include "params.cuh"
#include "kernels.h"

namespace abopt {

using namespace tc;

constexpr int AL_BM = 128, AL_BN = 128, AL_K = 64;
constexpr int AL_BOX_BYTES = 128 * 32 * 4;             // one TMA box: 128 rows x 32 floats = 16 KB
constexpr int AL_OPER_BYTES = 2 * AL_BOX_BYTES;        // 128 rows x 64 floats (two boxes along K)
constexpr int AL_MAXCOLS = 512;                        // longest key axis (two CTAs x 256 TMEM columns)
constexpr int AL_A_OFF = 0;                            // query operand: hi | lo

struct AttnLogitsArgs {
  int L, Lp, b0;
  const float* QA;                       // [N][H][L][64] packed query operand (read row by row into tensor memory)
  const float* rq; const float* rk;      // [N][H][L]
  const float* bias;                     // [N][H][L queries][Lp] (key index contiguous, like alpha)
  const uint8_t* mask;                   // [N][L]
  float* alpha;                          // [chunk][H][L][Lp]
  const uint8_t* exclude;                // optional [N][L]: keys left out of the softmax altogether (context cache, k_pair.cu)
  float2* stats;                         // optional [N][H][L]: (row maximum, sum of exp(l - maximum)) of every query row
};

void attn_debug_clocks(long long* out16) { for (int i = 0; i < 10; ++i) out16[i] = 0; }      // (the timeline hook of the retired one-tile kernel)

// smem box -> global through a 3-D tensor map (c0 = key, c1 = query row, c2 = (complex, head)); bulk async group
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// ------------------------------------------------------------------------------------------ logits + softmax
// attn_logits_persist_kernel: the phases of consecutive tiles overlap and no thread ever waits on a global load.  One CTA per SM walks the (complex, head, 128-query tile) list;
// 20 warps:
//   warp 0      TMA producer, one thread running a small event loop over two independent streams:
//                 operands -- key groups of 64 residues through a 2-stage ring;
//                 bias     -- the tile's pair bias in chunks of [128 queries][32 keys] (16 KB, swizzled) through a 6-slot ring,
//                             running ahead of the epilogue (the bias is the HBM read stream of this kernel)
//   warp 1      MMA issuer -- TS mode: the QUERY operand is read from tensor memory (2 slots of tf32 hi 64 | lo 64 columns, written
//               by the epilogue threads), the key operand from shared memory; one 256-column accumulator, released as soon as the
//               epilogue has moved it to registers, so tile n + 1 accumulates while the epilogue is busy with tile n; per key group
//               16 correction products first, then the 8 hi*hi ones (truncation note above)
//   warps 18-19 splitters  -- build the tf32 "lo" plane (x - trunc_tf32(x)) of every landed key group in shared memory,
//               so QA_lo / KB_lo never exist in global memory (saves their write in the projection kernel and their read here)
//   warps 2-17  epilogue   -- 4 threads per query row; thread (row, kq) owns keys 32 m + 8 kq + (0..7) of every chunk m:
//               TMEM -> registers (accumulator released at once), per chunk bias from shared memory (two conflict-free LDS.128
//               of the swizzled box) -> logits; max / exp / sum exchanged through shared memory; alpha leaves through swizzled
//               staging boxes and TMA tensor stores; the same threads load the query operand of tile n + 2 (16 floats each)
//               before the store phase and store it to TMEM after it
constexpr int AP_THREADS = 640, AP_EPI = 512, AP_SPLIT = 64;      // producer, MMA issuer, 16 epilogue warps, 2 splitter warps
// Shared memory (214 KB): two key-group stages (64 KB), a 6-deep pair-bias ring (96 KB), 3 alpha staging boxes (48 KB),
// double-buffered per-key tables, barriers.  Measured on B200 (round 2, scripts/kbench.py with variant builds, us per full-batch
// launch).  With the query operand in shared memory (64 KB): stages / bias slots / staging boxes = 2/3/2: 123.8, 1/4/3: 131.7 (the
// epilogue then waits for the MMA stream), 2/2/3: 131.1.  With it in tensor memory: 2/7/2: 123.3, 2/5/4: 123.4, 2/6/3: 120.0;
// staggering the odd CTAs by 3 or 5 us (to break the lockstep of the read-only bias phases and write-only store phases across
// the SMs): 120.  mbarrier.test_wait instead of try_wait in the producer's event loop: no change.  Neither ring depth nor phase
// alignment is what bounds it: DRAM is idle about half of the time (dram__cycles_active), the kernel is bound by the serial
// bias -> softmax -> store phases of a tile in its 16 epilogue warps.
#ifndef ABOPT_AP_BST
#define ABOPT_AP_BST 2
#endif
#ifndef ABOPT_AP_NBIAS
#define ABOPT_AP_NBIAS 6
#endif
#ifndef ABOPT_AP_NSTG
#define ABOPT_AP_NSTG 3
#endif
constexpr int AP_BST = ABOPT_AP_BST, AP_BGRP_BYTES = 4 * 64 * 32 * 4; // key-group stage: 64 keys x (hi k0 | hi k1 | lo k0 | lo k1) = 32 KB
constexpr int AP_KBOX = 64 * 32 * 4;                                  // one key box: 64 rows x 32 floats
constexpr int AP_NBIAS = ABOPT_AP_NBIAS, AP_BIAS_BYTES = 128 * 32 * 4;    // bias chunk: 128 queries x 32 keys
constexpr int AP_NSTG = ABOPT_AP_NSTG, AP_STG_BYTES = 128 * 32 * 4;       // alpha staging boxes of [128 queries][32 keys], 128-byte swizzle
constexpr int AP_B_OFF = 0;                                           // (the query operand lives in tensor memory)
constexpr uint32_t AP_TM_Q = 0, AP_TM_ACC = 256;                      // TMEM columns: 2 x (query hi 64 | lo 64) | accumulator (<= 256)
constexpr int AP_BIAS_OFF = AP_B_OFF + AP_BST * AP_BGRP_BYTES;
constexpr int AP_STG_OFF = AP_BIAS_OFF + AP_NBIAS * AP_BIAS_BYTES;
constexpr int AP_TAB_OFF = AP_STG_OFF + AP_NSTG * AP_STG_BYTES;       // ck[2][256] | pen[2][256] | xmax[4][128] | xsum[4][128]
constexpr int AP_BAR_OFF = AP_TAB_OFF + 4 * 256 * 4 + 8 * 128 * 4;
constexpr int AP_SMEM = AP_BAR_OFF + 256 + 4 * 128 * 4 + 1024;         // barriers (<= 31 x 8 B + slot) | SPLIT exchange [2][2][128] floats
static_assert(AP_SMEM <= 227 * 1024, "attn_logits_persist_kernel: shared memory");

// D[tmem] (+)= A[tmem] * B[smem]; one thread issues ("TS" mode: the A operand is read from tensor memory, lane = row, column = k)
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// registers -> 32 lanes x 16 consecutive fp32 columns (thread = lane); the caller waits (tcgen05.wait::st) once for a batch
__device__ __forceinline__ void tmem_st16_nowait(uint32_t taddr, const float (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                 "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                 "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
               : "memory");
}
// 2^x, x <= 0 (MUFU.EX2, 2 ulp; results below the normal range flush to zero -- attention weights < 1e-38)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void epi_sync512() { asm volatile("bar.sync 1, %0;" ::"n"(AP_EPI) : "memory"); }
// 32 lanes x 8 consecutive fp32 columns, no wait (the caller issues one tcgen05.wait::ld for a batch)
__device__ __forceinline__ void tmem_ld_32x8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}

struct AttnPersistArgs {
  const int2* windows;            // optional (focus mode): query windows (complex, first query row), 128 rows each; else the regular grid
  const int* wcount;              // device: wcount[1] = number of windows
  int nb_complex;                 // complexes covered by this launch
};

// tile index -> (complex in the launch, head, first query row): regular 128-row grid, or the focus windows (heads fastest)
struct TileRef { int bl, h, i0; };
__device__ __forceinline__ TileRef tile_ref(int tile, int nit, const int2* windows) {
  TileRef t;
  if (windows) { const int2 w = windows[tile / H]; t.bl = w.x; t.i0 = w.y; t.h = tile % H; }
  else { const int bh = tile / nit; t.i0 = (tile % nit) * 128; t.h = bh % H; t.bl = bh / H; }
  return t;
}

// non-blocking phase test for the producer's event loop (try_wait may suspend the thread for a system-dependent time slice)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
#ifdef ABOPT_AP_TESTWAIT
#define AP_POLL mbar_test_wait
#else
#define AP_POLL mbar_try_wait
#endif

// ---- thread-block-cluster helpers (key-split variant: two CTAs share one tile's softmax)
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `p` (own shared memory) inside CTA `rank` of the cluster, as a shared::cluster address
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t raddr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t raddr) {      // release at cluster scope: the store above is visible to the waiter
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t it = 0; !ok; ++it) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (it > 50000000u) __trap();
  }
}

// NCH: 32-key chunks per CTA and row (4: keys <= 128, 8: keys <= 256; SPLIT: 5..8 per half).
// SPLIT (256 < keys <= 512): the kernel runs as clusters of TWO CTAs; CTA `rank` of a cluster handles keys [rank * 32 NCH, ...) of
// the cluster's tile with exactly the single-CTA pipeline, and the two halves of a row exchange their maximum and their sum of
// exponentials through distributed shared memory (one remote store + one remote mbarrier arrive per row and quantity).
template <int NCH, bool SPLIT>
__global__ void __launch_bounds__(AP_THREADS, 1)
attn_logits_persist_kernel(const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
                           const __grid_constant__ CUtensorMap tmKh, const __grid_constant__ CUtensorMap tmKl,
                           const __grid_constant__ CUtensorMap tmBias, const __grid_constant__ CUtensorMap tmAl,
                           const AttnLogitsArgs a, const AttnPersistArgs pa) {
  constexpr int NGRP = (NCH + 1) / 2;                   // MMA groups of 64 keys (the last one has 32 when NCH is odd)
  constexpr int NLG = NCH * 8;                          // logits per epilogue thread
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* ck_tab = reinterpret_cast<float*>(smem + AP_TAB_OFF);      // [2][256]  rk of the tile's keys (double-buffered over tiles)
  float* pen_tab = ck_tab + 2 * 256;                                 // [2][256]  mask penalty of the tile's keys
  float* xmax = pen_tab + 2 * 256;        // [4][128]
  float* xsum = xmax + 4 * 128;           // [4][128]
  uint64_t* b_full = reinterpret_cast<uint64_t*>(smem + AP_BAR_OFF);      // [AP_BST]
  uint64_t* b_empty = b_full + AP_BST;    // [AP_BST]
  uint64_t* bias_full = b_empty + AP_BST;       // [AP_NBIAS]
  uint64_t* bias_empty = bias_full + AP_NBIAS;  // [AP_NBIAS]
  uint64_t* tmem_full = bias_empty + AP_NBIAS;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;   // [2]
  uint64_t* q_ready = tmem_empty + 2;     // [2]  the query operand of tile n is in TMEM slot n & 1
  uint64_t* b_split = q_ready + 2;        // [AP_BST]
  uint64_t* xch_bar = b_split + AP_BST;   // [2 kinds][2 tile parities]  SPLIT: the partner's row maxima / sums have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xch_bar + 4);
  float* xch = reinterpret_cast<float*>(smem + AP_BAR_OFF + 256);      // [2 kinds][2 parities][128 rows], written by the partner CTA

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = a.L, Lp = a.Lp;
  const int nit = (L + AL_BM - 1) / AL_BM;              // query tiles per (complex, head)
  const int ntiles = pa.windows ? pa.wcount[1] * H : pa.nb_complex * H * nit;
  const uint32_t crank = SPLIT ? cluster_ctarank() : 0u;
  const int key0 = (int)crank * NCH * 32;               // first key of this CTA
  const int tfirst = SPLIT ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // the tile walk of this CTA (SPLIT: of its cluster)
  const int tstride = SPLIT ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (threadIdx.x == 0) {
    mbar_init(&q_ready[0], AP_EPI / 32); mbar_init(&q_ready[1], AP_EPI / 32);
    for (int s = 0; s < AP_BST; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); mbar_init(&b_split[s], AP_SPLIT); }
    for (int s = 0; s < AP_NBIAS; ++s) { mbar_init(&bias_full[s], 1); mbar_init(&bias_empty[s], AP_EPI / 32); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], AP_EPI / 32); }
    for (int s = 0; s < 4; ++s) mbar_init(&xch_bar[s], 128);
    mbar_fence_init();
    tma_prefetch_desc(&tmQh); tma_prefetch_desc(&tmQl); tma_prefetch_desc(&tmKh); tma_prefetch_desc(&tmKl); tma_prefetch_desc(&tmBias);
    tma_prefetch_desc(&tmAl);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  if constexpr (SPLIT) cluster_sync_all();              // the partner's barriers are initialised before anyone arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: event loop over the operand stream and the bias stream =====================
    if (elect_one()) {
      int otile = tfirst, on = 0, og = 0, ostep = 0;        // operand stream: tile, tile count, key-group count, step in tile
      int btile = tfirst, bc = 0, bm = 0;                   // bias stream: tile, chunk count, chunk in tile
      while (otile < ntiles || btile < ntiles) {
        if (otile < ntiles) {
          const TileRef tr = tile_ref(otile, nit, pa.windows);
          const int row_base = ((a.b0 + tr.bl) * H + tr.h) * L;
          if (ostep == 0) {
            ostep = 1;                                      // (the query operand is loaded by the splitter warps, into TMEM)
          } else {
            const int s = og % AP_BST;
            if (AP_POLL(&b_empty[s], ((og / AP_BST) & 1) ^ 1)) {
              unsigned char* B = smem + AP_B_OFF + s * AP_BGRP_BYTES;
              const int r0 = row_base + key0 + (ostep - 1) * 64;
              mbar_expect_tx(&b_full[s], 2 * AP_KBOX);
              tma_load_2d(B, &tmKh, 0, r0, &b_full[s]);
              tma_load_2d(B + AP_KBOX, &tmKh, 32, r0, &b_full[s]);
              ++og;
              if (++ostep > NGRP) { ostep = 0; ++on; otile += tstride; }
            }
          }
        }
        if (btile < ntiles) {
          const int s = bc % AP_NBIAS;
          if (AP_POLL(&bias_empty[s], ((bc / AP_NBIAS) & 1) ^ 1)) {
            const TileRef tr = tile_ref(btile, nit, pa.windows);
            const int row_base = ((a.b0 + tr.bl) * H + tr.h) * L;
            mbar_expect_tx(&bias_full[s], AP_BIAS_BYTES);
            tma_load_2d(smem + AP_BIAS_OFF + s * AP_BIAS_BYTES, &tmBias, key0 + bm * 32, row_base + tr.i0, &bias_full[s]);
            ++bc;
            if (++bm == NCH) { bm = 0; btile += tstride; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc64 = idesc_tf32(AL_BM, 64), idesc32 = idesc_tf32(AL_BM, 32);
    int n = 0, g = 0;
    for (int tile = tfirst; tile < ntiles; tile += tstride, ++n) {
      mbar_wait(&tmem_empty[0], (n & 1) ^ 1);             // the epilogue has moved the accumulator of tile n - 1 to registers
      mbar_wait(&q_ready[n & 1], (n >> 1) & 1);           // query operand hi | lo of this tile are in TMEM slot n & 1
      for (int gi = 0; gi < NGRP; ++gi, ++g) {
        const int s = g % AP_BST;
        mbar_wait(&b_split[s], (g / AP_BST) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_hi = tmem_base + AP_TM_Q + (n & 1) * 128, a_lo = a_hi + 64;      // TS mode: A from tensor memory, column = k
          const uint32_t b_hi = smem_u32(smem + AP_B_OFF + s * AP_BGRP_BYTES), b_lo = b_hi + 2 * AP_KBOX;
          const uint32_t d = tmem_base + AP_TM_ACC + gi * 64;
          const uint32_t idesc = ((NCH & 1) && gi == NGRP - 1) ? idesc32 : idesc64;      // 32-key tail group (its stage holds 64 rows)
#pragma unroll
          for (int kk = 0; kk < AL_K / 8; ++kk) {
            const uint32_t kb = (kk >> 2) * AP_KBOX + (kk & 3) * 32;
            mma_tf32_ts(d, a_hi + kk * 8, smem_desc_sw128(b_lo + kb), idesc, kk == 0 ? 0u : 1u);
            mma_tf32_ts(d, a_lo + kk * 8, smem_desc_sw128(b_hi + kb), idesc, 1u);
          }
#pragma unroll
          for (int kk = 0; kk < AL_K / 8; ++kk) {
            const uint32_t kb = (kk >> 2) * AP_KBOX + (kk & 3) * 32;
            mma_tf32_ts(d, a_hi + kk * 8, smem_desc_sw128(b_hi + kb), idesc, 1u);
          }
          mma_commit(&b_empty[s]);
          if (gi == NGRP - 1) mma_commit(&tmem_full[0]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 18) {
    // ===================== splitters (warps 18-19): lo = x - trunc_tf32(x) of every landed key group =====================
    const int st = threadIdx.x - 18 * 32;                 // 0..63
    auto split = [&](const unsigned char* hi, unsigned char* lo, int n16) {
      const float4* src = reinterpret_cast<const float4*>(hi);
      float4* dst = reinterpret_cast<float4*>(lo);
#pragma unroll 4
      for (int k = st; k < n16; k += AP_SPLIT) {
        const float4 v = src[k];
        dst[k] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
      }
      fence_async_smem();                                 // generic-proxy writes -> visible to the tensor core (async proxy)
    };
    int g = 0;
    for (int tile = tfirst; tile < ntiles; tile += tstride)
      for (int gi = 0; gi < NGRP; ++gi, ++g) {
        const int s = g % AP_BST;
        mbar_wait(&b_full[s], (g / AP_BST) & 1);
        unsigned char* B = smem + AP_B_OFF + s * AP_BGRP_BYTES;
        split(B, B + 2 * AP_KBOX, 2 * AP_KBOX / 16);
        mbar_arrive(&b_split[s]);
      }
  } else {
    // ===================== epilogue (warps 2..17) =====================
    const int q = warp & 3;                               // TMEM lane quarter this warp may access
    const int kq = (warp - 2) >> 2;                       // which 8 keys of every 32-key chunk this thread handles
    const int te = q * 32 + lane;                         // 0..127: query row in the tile
    const int et = (warp - 2) * 32 + lane;                // 0..511
    const float scale = 0.57735026918962576f;             // sqrt(1/3), ga.py:166
    const float l2e = 1.4426950408889634f;
    // per-key tables of a tile (rk[j], and the mask penalty: 1e5 for masked keys, +inf beyond the end -> alpha = 0) and the
    // row's rq: loaded one tile AHEAD into registers (threads et < 32 NCH own one key each), parked in the other table buffer at
    // the end of the tile, so that no tile starts by waiting on global loads
    auto key_tables = [&](const TileRef& tr, float& ckv, float& penv, float& rqv) {
      const int b = a.b0 + tr.bl, row_base = (b * H + tr.h) * L;
      if (et < NCH * 32) {
        const int j = key0 + et;
        ckv = (j < L) ? __ldg(a.rk + (size_t)row_base + j) : 0.f;
        penv = (j < L) ? (a.mask[(size_t)b * L + j] != 0 ? 0.f : 1e5f) : INFINITY;
        if (a.exclude && j < L && a.exclude[(size_t)b * L + j] != 0) penv = INFINITY;
      }
      rqv = (tr.i0 + te < L) ? __ldg(a.rq + (size_t)row_base + tr.i0 + te) : 0.f;
    };
    int n = 0, bc = 0, sc = 0;
    float rqi = 0.f;
    // The query operand lives in TENSOR MEMORY (TS-mode MMAs; its 64 KB of shared memory went to the pair-bias ring): thread
    // (row, kq) owns 16 of the row's 64 floats, reads them from global memory and stores tf32 hi | lo to columns 16 kq.. of the
    // hi / lo halves of TMEM slot (tile count & 1).  The loads of tile n + 2 are issued before the store phase of tile n and
    // stored after it (slot n & 1 is free: the MMAs of tile n completed before this epilogue started).
    auto q_load = [&](int tile, float4 (&qv)[4]) {
      const TileRef tr = tile_ref(tile, nit, pa.windows);
      const int qi = tr.i0 + te;
      const float4* qrow = reinterpret_cast<const float4*>(a.QA + ((size_t)((a.b0 + tr.bl) * H + tr.h) * L + (qi < L ? qi : 0)) * 64) + kq * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e) qv[e] = qi < L ? __ldg(qrow + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto q_store = [&](int slot, const float4 (&qv)[4]) {
      const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + AP_TM_Q + slot * 128 + kq * 16;
      float hi[16], lo[16];
#pragma unroll
      for (int e = 0; e < 4; ++e) { hi[4 * e] = qv[e].x; hi[4 * e + 1] = qv[e].y; hi[4 * e + 2] = qv[e].z; hi[4 * e + 3] = qv[e].w; }
#pragma unroll
      for (int e = 0; e < 16; ++e) lo[e] = tf32_lo(hi[e]);
      tmem_st16_nowait(tq, hi);                           // (raw fp32: the tensor core ignores the low 13 mantissa bits)
      tmem_st16_nowait(tq + 64, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_ready[slot]);
    };
    if (tfirst < ntiles) {
      float ckv = 0.f, penv = 0.f;
      key_tables(tile_ref(tfirst, nit, pa.windows), ckv, penv, rqi);
      if (et < NCH * 32) { ck_tab[et] = ckv; pen_tab[et] = penv; }
      float4 qv[4];
      q_load(tfirst, qv); q_store(0, qv);
      if (tfirst + tstride < ntiles) { q_load(tfirst + tstride, qv); q_store(1, qv); }
    }
    for (int tile = tfirst; tile < ntiles; tile += tstride, ++n) {
      const TileRef tr = tile_ref(tile, nit, pa.windows);
      const int h = tr.h, bl = tr.bl, i0 = tr.i0;
      const int buf = n & 1;
      const float* ck = ck_tab + buf * 256;
      const float* pen = pen_tab + buf * 256;
      epi_sync512();                                      // this tile's tables visible; also: every warp is done with the previous tile
      float ck_n = 0.f, pen_n = 0.f, rq_n = 0.f;
      const bool more = tile + tstride < ntiles;
      if (more) key_tables(tile_ref(tile + tstride, nit, pa.windows), ck_n, pen_n, rq_n);      // in flight during this tile
      mbar_wait(&tmem_full[0], n & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + AP_TM_ACC + kq * 8;
      float lg[NLG];
      {
        uint32_t raw[NCH][8];
#pragma unroll
        for (int m = 0; m < NCH; ++m) tmem_ld_32x8_nowait(trow + m * 32, raw[m]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int m = 0; m < NCH; ++m)
#pragma unroll
          for (int e = 0; e < 8; ++e) lg[m * 8 + e] = __uint_as_float(raw[m][e]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[0]);         // the accumulator is in registers: the MMA warp may start the next tile
      float mx = -INFINITY;
#pragma unroll
      for (int m = 0; m < NCH; ++m, ++bc) {
        const int s = bc % AP_NBIAS;
        mbar_wait(&bias_full[s], (bc / AP_NBIAS) & 1);
        // chunk = box of [128 queries][32 keys], 128-byte swizzle: row te, 16-byte chunks 2 kq and 2 kq + 1
        const unsigned char* bs = smem + AP_BIAS_OFF + s * AP_BIAS_BYTES + te * 128;
        const float4 b0v = *reinterpret_cast<const float4*>(bs + (((2 * kq) ^ (te & 7)) << 4));
        const float4 b1v = *reinterpret_cast<const float4*>(bs + (((2 * kq + 1) ^ (te & 7)) << 4));
        const float bv[8] = {b0v.x, b0v.y, b0v.z, b0v.w, b1v.x, b1v.y, b1v.z, b1v.w};
        const int j0 = m * 32 + kq * 8;
        const float4 c0 = *reinterpret_cast<const float4*>(ck + j0), c1 = *reinterpret_cast<const float4*>(ck + j0 + 4);
        const float4 p0 = *reinterpret_cast<const float4*>(pen + j0), p1 = *reinterpret_cast<const float4*>(pen + j0 + 4);
        const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        const float pp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float lgt = ((lg[m * 8 + e] + bv[e]) + (rqi + cc[e])) * scale - pp[e];
          mx = fmaxf(mx, lgt);
          lg[m * 8 + e] = lgt;
        }
        // Release the slot only AFTER the bias values have been consumed.  An arrive issued right behind the two LDS does not
        // wait for their data (the scoreboard wait sits at the first use), and mbarrier operations are not ordered behind
        // queued shared-memory loads: the producer's refill of this slot was observed to overtake the loads of the last
        // arriving warp (round 2: ~0.5 % of the tiles of a full C2 batch came out with one stale 16-byte bias piece per row).
        __syncwarp();
        if (lane == 0) mbar_arrive(&bias_empty[s]);
      }
      xmax[kq * 128 + te] = mx;
      epi_sync512();
      mx = fmaxf(fmaxf(xmax[te], xmax[128 + te]), fmaxf(xmax[256 + te], xmax[384 + te]));
      if constexpr (SPLIT) {
        // the row maximum over BOTH key halves: send ours to the partner CTA, take the partner's (buffer / barrier phase by tile
        // parity: the partner cannot be more than one tile ahead, it needs our value of the tile in between)
        if (kq == 0) {
          st_cluster_f32(mapa_u32(&xch[(0 * 2 + buf) * 128 + te], crank ^ 1u), mx);
          mbar_arrive_cluster(mapa_u32(&xch_bar[0 * 2 + buf], crank ^ 1u));
        }
        mbar_wait_cluster(&xch_bar[0 * 2 + buf], (n >> 1) & 1);
        mx = fmaxf(mx, xch[(0 * 2 + buf) * 128 + te]);
      }
      if (a.exclude && mx == -INFINITY) mx = 0.f;         // every key excluded: alpha = 0 (and stats = (0, 0)), not NaN
      // exp(l - m) = 2^((l - m) log2e): subtract FIRST (exact near the maximum, where the attention mass is)
      float sum = 0.f;
#pragma unroll
      for (int e = 0; e < NLG; ++e) { lg[e] = ex2_approx((lg[e] - mx) * l2e); sum += lg[e]; }
      xsum[kq * 128 + te] = sum;
      // the next tile's tables: the other buffer was last read in the bias phase of the previous tile
      if (more && et < NCH * 32) { ck_tab[(buf ^ 1) * 256 + et] = ck_n; pen_tab[(buf ^ 1) * 256 + et] = pen_n; }
      rqi = rq_n;
      epi_sync512();
      float rowsum = (xsum[te] + xsum[128 + te]) + (xsum[256 + te] + xsum[384 + te]);
      if constexpr (SPLIT) {
        if (kq == 0) {
          st_cluster_f32(mapa_u32(&xch[(1 * 2 + buf) * 128 + te], crank ^ 1u), rowsum);
          mbar_arrive_cluster(mapa_u32(&xch_bar[1 * 2 + buf], crank ^ 1u));
        }
        mbar_wait_cluster(&xch_bar[1 * 2 + buf], (n >> 1) & 1);
        const float other = xch[(1 * 2 + buf) * 128 + te];
        rowsum = crank == 0 ? rowsum + other : other + rowsum;      // same order of the two halves in both CTAs
      }
      const float inv = (a.exclude && rowsum == 0.f) ? 0.f : 1.0f / rowsum;
      if (a.stats && kq == 0 && crank == 0 && i0 + te < L)
        a.stats[(size_t)((a.b0 + bl) * H + h) * L + i0 + te] = make_float2(mx, rowsum);
      float4 qnext[4];
      const bool more2 = tile + 2 * tstride < ntiles;
      if (more2) q_load(tile + 2 * tstride, qnext);       // in flight during the store phase
      // alpha leaves chunk by chunk through AP_NSTG staging boxes ([128 queries][32 keys], 128-byte swizzle) and TMA tensor
      // stores: whole lines, rows >= L and keys >= Lp clipped by the hardware.  (Per-lane 256-bit global stores of a
      // row-per-lane layout cost 32 sector requests per instruction and kept the LSU the bottleneck of this kernel.)
      // Box sc % AP_NSTG was last read by the store of chunk sc - AP_NSTG; thread 0 waited for every store but the AP_NSTG - 2
      // most recent ones BEFORE the previous barrier, so the box is free.
#pragma unroll
      for (int m = 0; m < NCH; ++m, ++sc) {
        unsigned char* stg = smem + AP_STG_OFF + (sc % AP_NSTG) * AP_STG_BYTES;
        unsigned char* rowp = stg + te * 128;
        *reinterpret_cast<float4*>(rowp + (((2 * kq) ^ (te & 7)) << 4)) =
            make_float4(lg[m * 8] * inv, lg[m * 8 + 1] * inv, lg[m * 8 + 2] * inv, lg[m * 8 + 3] * inv);
        *reinterpret_cast<float4*>(rowp + (((2 * kq + 1) ^ (te & 7)) << 4)) =
            make_float4(lg[m * 8 + 4] * inv, lg[m * 8 + 5] * inv, lg[m * 8 + 6] * inv, lg[m * 8 + 7] * inv);
        fence_async_smem();
        if (et == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(AP_NSTG - 2) : "memory");
        epi_sync512();
        if (et == 0) {                                     // (an empty group when the chunk lies beyond Lp keeps the count in step)
          if (key0 + m * 32 < Lp) tma_store_3d(&tmAl, stg, key0 + m * 32, i0, bl * H + h);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (more2) q_store(n & 1, qnext);
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");          // smem must outlive the reads
    tc_fence_before();
  }
  __syncthreads();
  if constexpr (SPLIT) cluster_sync_all();              // the partner may still be writing our exchange buffers
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

bool attn_needs_qk_lo(int) { return false; }      // the lo planes of QA / KB are built on chip for every length
cudaError_t attn_tc_init() {
  cudaError_t e = cudaFuncSetAttribute(attn_logits_persist_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AP_SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(attn_logits_persist_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AP_SMEM);
  if (e != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(attn_logits_persist_kernel<5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AP_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(attn_logits_persist_kernel<6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AP_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(attn_logits_persist_kernel<7, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AP_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(attn_logits_persist_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AP_SMEM)) != cudaSuccess) return e;
  return cudaSuccess;
}

bool launch_attn_logits_tc(int nb, int b0, int N, int L, int Lp, const AttnOperands& op, const float* bias_layer, const uint8_t* mask,
                           float* alpha, cudaStream_t st, const int2* windows, const int* wcount, const uint8_t* exclude, float2* stats) {
  if (L > AL_MAXCOLS) return false;
  CUtensorMap qh, kh64, bm32, al;
  const uint64_t rows = (uint64_t)N * H * L;
  // query operand boxes of 128 rows, key operands in groups of 64 residues (raw fp32 = the "hi" plane; the lo planes are built on
  // chip); the pair bias [(b,h,i) rows][Lp keys] as swizzled [128 queries][32 keys] boxes
  if (!make_tmap(&qh, op.QA, rows, 64, 64, 128) || !make_tmap(&kh64, op.KB, rows, 64, 64, 64) ||
      !make_tmap(&bm32, bias_layer, rows, (uint64_t)Lp, (uint64_t)Lp, 128))
    return false;
  // alpha as a 3-D tensor [chunk * H][L queries][Lp keys] for the TMA stores (rows >= L are clipped)
  if (!make_tmap_3d(&al, alpha, Lp, L, (uint64_t)nb * H, 32, 128)) return false;
  ProfScope prof__(KK_LOGITS, st);
  AttnLogitsArgs a{L, Lp, b0, op.QA, op.rq, op.rk, bias_layer, mask, alpha, exclude, stats};
  const int ncols = ((L + AL_BN - 1) / AL_BN) * AL_BN;
  const int ntiles = nb * H * ((L + AL_BM - 1) / AL_BM);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const AttnPersistArgs pa{windows, wcount, nb};
  if (ncols <= 256) {
    if (ncols <= 128) attn_logits_persist_kernel<4, false><<<ntiles < sms ? ntiles : sms, AP_THREADS, AP_SMEM, st>>>(qh, qh, kh64, kh64, bm32, al, a, pa);
    else attn_logits_persist_kernel<8, false><<<ntiles < sms ? ntiles : sms, AP_THREADS, AP_SMEM, st>>>(qh, qh, kh64, kh64, bm32, al, a, pa);
    return true;
  }
  // 256 < keys <= 512: clusters of two CTAs, each takes half of the keys of a tile (NCH 32-key chunks)
  const int nch = ((Lp + 31) / 32 + 1) / 2;
  const int nclusters = ntiles < sms / 2 ? ntiles : sms / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * nclusters); cfg.blockDim = dim3(AP_THREADS); cfg.dynamicSmemBytes = AP_SMEM; cfg.stream = st;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  cudaError_t e;
  switch (nch) {
    case 5: e = cudaLaunchKernelEx(&cfg, attn_logits_persist_kernel<5, true>, qh, qh, kh64, kh64, bm32, al, a, pa); break;
    case 6: e = cudaLaunchKernelEx(&cfg, attn_logits_persist_kernel<6, true>, qh, qh, kh64, kh64, bm32, al, a, pa); break;
    case 7: e = cudaLaunchKernelEx(&cfg, attn_logits_persist_kernel<7, true>, qh, qh, kh64, kh64, bm32, al, a, pa); break;
    default: e = cudaLaunchKernelEx(&cfg, attn_logits_persist_kernel<8, true>, qh, qh, kh64, kh64, bm32, al, a, pa); break;
  }
  return e == cudaSuccess;
}

// ------------------------------------------------------------------------------------------ aggregation GEMM
// Node and point aggregation  O[i][n] = sum_j alpha[i][j] V[j][n]  (ga.py:120-136) on the tensor cores, followed by the
// local-frame features (ga.py:137-146).  V^T[b][h][n][j] = [ value channels (32) | global value points (24) | 0 (8) ] comes
// K-major (key index contiguous) from the projection epilogue, alpha from the logits kernel.  Per key block of 32: alpha 128 x 32
// and V^T 64 x 32 arrive raw (fp32 = the tf32 "hi" plane); hi*hi goes to one of two 64-column TMEM accumulators (one per half of
// the keys), hi*lo + lo*hi to a third (short accumulation chains: the tensor core truncates the fp32 accumulator on every
// accumulation, see k_tc.cu).
// The A operand (alpha) is fed from TENSOR MEMORY (TS-mode MMAs): the splitter threads (thread = query row) move each landed alpha
// box from shared memory into a 2-slot TMEM ring as hi | lo planes, so the MMAs read only V^T from shared memory (2 KB instead of
// 6 KB per instruction) and a stage of the shared-memory ring shrinks from 48 to 32 KB (6 stages in flight instead of 4).
constexpr int AG2_A_BYTES = 128 * 32 * 4, AG2_B_BYTES = 64 * 32 * 4;
constexpr int AG2_STAGE_BYTES = AG2_A_BYTES + 2 * AG2_B_BYTES;          // 32 KB: alpha raw | V^T hi | V^T lo
constexpr int AG2_TX_BYTES = AG2_A_BYTES + AG2_B_BYTES;                 // what TMA delivers per stage (the lo planes are built on chip)
constexpr uint32_t AG2_TM_A = 0, AG2_TM_ACC = 128, AG2_ACC_COLS = 192;  // TMEM: 2 x (alpha hi 32 | lo 32) | 2 x (main 64 | main 64 | corrections 64)

struct AggrArgs {
  int L, Lp, b0;
  const float* R; const float* t;        // [N][L][3][3], [N][L][3]
  float* feat;                           // [N][L][1824]
};

// ------------------------------------------------------------------------------------------ persistent aggregation GEMM
// aggr_persist_kernel: structured like attn_logits_persist_kernel.  One CTA per SM walks
// the (complex, head, 128-query tile) list; the k-block ring runs across tile boundaries and the epilogue of tile n overlaps
// the main loop of tile n + 1 (two TMEM accumulator sets):
//   warp 0      TMA producer: key blocks of 32 through a 6-stage ring (alpha 128 x 32 and V^T 64 x 32, raw fp32 = the hi planes)
//   warp 1      MMA issuer (waits for "split"), A from tensor memory
//   warps 2-5   splitters, thread = query row: alpha box -> TMEM hi | lo (tcgen05.st); tf32 lo plane of the V^T box in shared
//               memory (no VT_lo in global memory: 50 MB per layer less to write in the projection kernel and to read here)
//   warps 6-9   epilogue: thread = query row: 64 sums -> node aggregate, R_i^T (o - t_i), norms, directions; every feature
//               group of a row is a 32-byte-aligned run, stored with 256-bit stores (full sectors, no staging buffer)
constexpr int AGP_THREADS = 320, AGP_ST = 6;
// Tile walk from the FIRST complex to the last: pair_stream_kernel has just read alpha from the last row to the first, so what is
// still in the 126 MB L2 is the head of it -- and this kernel is bound by the latency of its ring, not by bandwidth.
#define AGP_TILE(t) (t)
constexpr int AGP_SMEM = AGP_ST * AG2_STAGE_BYTES + 256 + 1024;

__device__ __forceinline__ void st_v8f(float* p, float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a0), "f"(a1), "f"(a2), "f"(a3), "f"(a4),
               "f"(a5), "f"(a6), "f"(a7) : "memory");
}
// registers -> 32 lanes x 32 consecutive fp32 columns (thread = lane); the caller waits (tcgen05.wait::st) once for a batch
__device__ __forceinline__ void tmem_st32_nowait(uint32_t taddr, const float (&v)[32]) {
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

__global__ void __launch_bounds__(AGP_THREADS, 1)
aggr_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmVh,
                    const AggrArgs a, const int nb_complex,
                    const int2* windows, const int* wcount, const int* cidx) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + AGP_ST * AG2_STAGE_BYTES);
  uint64_t* split = full + AGP_ST;
  uint64_t* empty = split + AGP_ST;
  uint64_t* tmem_full = empty + AGP_ST;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint64_t* ta_free = tmem_empty + 2;        // [2]  the MMAs that read TMEM alpha slot s have completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ta_free + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = a.L;
  const int nkb = (L + 31) / 32;
  const int gsz = (nkb + 1) / 2;                        // key blocks per main accumulator (two short accumulation chains)
  const int nit = (L + 127) / 128;
  const int ntiles = windows ? wcount[1] * H : nb_complex * H * nit;

  if (threadIdx.x == 0) {
    for (int s = 0; s < AGP_ST; ++s) { mbar_init(&full[s], 1); mbar_init(&split[s], 4); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); mbar_init(&ta_free[b], 1); }
    mbar_fence_init();
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmVh);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int g = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const TileRef tr = tile_ref(AGP_TILE(tile), nit, windows);
        const int arow = (tr.bl * H + tr.h) * L + tr.i0, vrow = ((a.b0 + tr.bl) * H + tr.h) * 64;
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int s = g % AGP_ST;
          mbar_wait(&empty[s], ((g / AGP_ST) & 1) ^ 1);
          unsigned char* st = smem + s * AG2_STAGE_BYTES;
          mbar_expect_tx(&full[s], AG2_TX_BYTES);
          tma_load_2d(st, &tmA, kb * 32, arow, &full[s]);
          tma_load_2d(st + AG2_A_BYTES, &tmVh, kb * 32, vrow, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = idesc_tf32(128, 64);
    int g = 0, n = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++n) {
      const int buf = n & 1;
      mbar_wait(&tmem_empty[buf], ((n >> 1) & 1) ^ 1);
      const uint32_t tb = tmem_base + AG2_TM_ACC + buf * AG2_ACC_COLS;
      for (int kb = 0; kb < nkb; ++kb, ++g) {
        const int s = g % AGP_ST;
        mbar_wait(&split[s], (g / AGP_ST) & 1);             // TMA data landed, alpha moved to TMEM, V^T lo plane built
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_hi = smem_u32(smem + s * AG2_STAGE_BYTES + AG2_A_BYTES), b_lo = b_hi + AG2_B_BYTES;
          const uint32_t ta = tmem_base + AG2_TM_A + (g & 1) * 64;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t ah = ta + k * 8, al = ah + 32;
            const uint64_t dbh = smem_desc_sw128(b_hi + k * 32), dbl = smem_desc_sw128(b_lo + k * 32);
            mma_tf32_ts(tb + (kb / gsz) * 64, ah, dbh, idesc, (kb % gsz == 0 && k == 0) ? 0u : 1u);
            mma_tf32_ts(tb + 128, ah, dbl, idesc, (kb == 0 && k == 0) ? 0u : 1u);
            mma_tf32_ts(tb + 128, al, dbh, idesc, 1u);
          }
          mma_commit(&empty[s]);
          mma_commit(&ta_free[g & 1]);
          if (kb == nkb - 1) mma_commit(&tmem_full[buf]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ---- splitters: thread = query row of the tile
    const int te = (warp - 2) * 32 + lane;                // 0..127: position in the V^T split
    const int row = (warp & 3) * 32 + lane;               // TMEM lane = query row this thread may write
    const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + AG2_TM_A;
    int g = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
      for (int kb = 0; kb < nkb; ++kb, ++g) {
        const int s = g % AGP_ST;
        mbar_wait(&full[s], (g / AGP_ST) & 1);
        // everything that needs only the landed stage first: V^T lo plane, this row of the alpha box and its lo plane
        const float4* vsrc = reinterpret_cast<const float4*>(smem + s * AG2_STAGE_BYTES + AG2_A_BYTES);
        float4* vdst = reinterpret_cast<float4*>(smem + s * AG2_STAGE_BYTES + AG2_A_BYTES + AG2_B_BYTES);
#pragma unroll
        for (int m = 0; m < AG2_B_BYTES / 16 / 128; ++m) {
          const float4 v = vsrc[te + 128 * m];
          vdst[te + 128 * m] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
        }
        // alpha box [128 queries][32 keys], 128-byte swizzle: row `row`, 16-byte unit u sits at u ^ (row & 7)
        const unsigned char* ar = smem + s * AG2_STAGE_BYTES + row * 128;
        float hi[32], lo[32];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4 v = *reinterpret_cast<const float4*>(ar + ((u ^ (row & 7)) << 4));
          hi[4 * u] = v.x; hi[4 * u + 1] = v.y; hi[4 * u + 2] = v.z; hi[4 * u + 3] = v.w;
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) lo[e] = tf32_lo(hi[e]);
        mbar_wait(&ta_free[g & 1], ((g >> 1) & 1) ^ 1);     // the MMAs of key block g - 2 have read this TMEM slot
        tc_fence_after();
        tmem_st32_nowait(trow + (g & 1) * 64, hi);          // (raw fp32: the tensor core ignores the low 13 mantissa bits)
        tmem_st32_nowait(trow + (g & 1) * 64 + 32, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&split[s]);
      }
  } else {
    // ---- epilogue
    const int q = warp & 3;
    const int ngrp = (nkb + gsz - 1) / gsz;
    int n = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++n) {
      const TileRef tr = tile_ref(AGP_TILE(tile), nit, windows);
      const int h = tr.h, b = a.b0 + tr.bl;
      const int buf = n & 1;
      const int i = tr.i0 + q * 32 + lane;
      mbar_wait(&tmem_full[buf], (n >> 1) & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + AG2_TM_ACC + buf * AG2_ACC_COLS;
      float o[64];
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        float v[32], w[32];
        tmem_ld_32x32(trow + 128 + c, w);                   // corrections
        tmem_ld_32x32(trow + c, v);
#pragma unroll
        for (int e = 0; e < 32; ++e) o[c + e] = v[e];
        if (ngrp > 1) {
          tmem_ld_32x32(trow + 64 + c, v);
#pragma unroll
          for (int e = 0; e < 32; ++e) o[c + e] += v[e];
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) o[c + e] += w[e];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);        // sums are in registers: the MMA warp may reuse the buffer
      const int crow = (i < L) ? (cidx ? cidx[(size_t)b * L + i] : b * L + i) : -1;      // focus mode: compact output row, -1 = not needed
      if (crow >= 0) {
        const size_t row = (size_t)b * L + i;
        float* fr = a.feat + (size_t)crow * NFEAT;
        // node aggregate (ga.py:120-125)
#pragma unroll
        for (int d = 0; d < D; d += 8) st_v8f(fr + FEAT_NODE + h * D + d, o[d], o[d + 1], o[d + 2], o[d + 3], o[d + 4], o[d + 5], o[d + 6], o[d + 7]);
        // point aggregate -> local frame p = R^T (q - t) (geometry.py:94-113), norm, direction (eps 1e-4, ga.py:139)
        float Rm[9], tv[3];
#pragma unroll
        for (int k = 0; k < 9; ++k) Rm[k] = __ldg(a.R + row * 9 + k);
#pragma unroll
        for (int k = 0; k < 3; ++k) tv[k] = __ldg(a.t + row * 3 + k);
        float pts[24], nrm[8], dir[24];
#pragma unroll
        for (int p = 0; p < P; ++p) {
          const float gx = o[D + p * 3 + 0] - tv[0], gy = o[D + p * 3 + 1] - tv[1], gz = o[D + p * 3 + 2] - tv[2];
          const float lx = Rm[0] * gx + Rm[3] * gy + Rm[6] * gz;
          const float ly = Rm[1] * gx + Rm[4] * gy + Rm[7] * gz;
          const float lz = Rm[2] * gx + Rm[5] * gy + Rm[8] * gz;
          const float nn = sqrtf(lx * lx + ly * ly + lz * lz);
          const float den = nn + 1e-4f;
          pts[p * 3] = lx; pts[p * 3 + 1] = ly; pts[p * 3 + 2] = lz;
          nrm[p] = nn;
          dir[p * 3] = lx / den; dir[p * 3 + 1] = ly / den; dir[p * 3 + 2] = lz / den;
        }
#pragma unroll
        for (int d = 0; d < 24; d += 8) {
          st_v8f(fr + FEAT_PTS + h * P * 3 + d, pts[d], pts[d + 1], pts[d + 2], pts[d + 3], pts[d + 4], pts[d + 5], pts[d + 6], pts[d + 7]);
          st_v8f(fr + FEAT_DIR + h * P * 3 + d, dir[d], dir[d + 1], dir[d + 2], dir[d + 3], dir[d + 4], dir[d + 5], dir[d + 6], dir[d + 7]);
        }
        st_v8f(fr + FEAT_DIST + h * P, nrm[0], nrm[1], nrm[2], nrm[3], nrm[4], nrm[5], nrm[6], nrm[7]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

cudaError_t aggr_tc_init() {
  return cudaFuncSetAttribute(aggr_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AGP_SMEM);
}

// alpha: [chunk][H][L][Lp]; VT: [N][H][64][Lp]
bool launch_aggr_tc(int nb, int b0, int N, int L, int Lp, const float* alpha, const float* VT,
                    const float* R, const float* t, float* feat, cudaStream_t st, const int2* windows, const int* wcount,
                    const int* cidx) {
  CUtensorMap ah, vh;
  const uint64_t arows = (uint64_t)nb * H * L, vrows = (uint64_t)N * H * 64;
  if (!make_tmap(&ah, alpha, arows, Lp, Lp, 128) || !make_tmap(&vh, VT, vrows, Lp, Lp, 64)) return false;
  ProfScope prof__(KK_AGGR, st);
  AggrArgs a{L, Lp, b0, R, t, feat};
  const int ntiles = nb * H * ((L + 127) / 128);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  aggr_persist_kernel<<<ntiles < sms ? ntiles : sms, AGP_THREADS, AGP_SMEM, st>>>(ah, vh, a, nb, windows, wcount, cidx);
  return true;
}

}  // namespace abopt
