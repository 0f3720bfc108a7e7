// outT_tail_kernel: the whole tail of a GABlock in ONE kernel on the 5th-gen tensor cores (ga.py:173-178):
//   out_transform (1824 -> 128, bias) -> mask_zero -> LayerNorm(x + .) -> 3-layer ReLU MLP -> LayerNorm(res + .)
// Phase 1 is gemm3x_splitA_kernel (k_tc.cu): D[128 rows][128] = feat[128][1824] * W_out^T as 3xTF32, raw feat streamed
// from HBM through a 3-stage TMA ring, its tf32 lo plane built on chip, accumulators promoted every 8 k-blocks.
// Phase 2 keeps the row tile on chip: the epilogue threads (thread = row x half of the 128 columns) apply bias / mask /
// residual / LayerNorm 1 in registers, write the activations as a K-major 128-byte-swizzled A operand (hi and lo plane)
// into the drained pipeline memory, and the three 128 x 128 MLP layers run as tcgen05 GEMMs whose weights (L2 resident)
// arrive through a 2-stage TMA ring; bias + ReLU happen between layers in registers, the residual and LayerNorm 2 at the
// end.  Replaces gemm3x + the FFMA tail_kernel (which spent 58 us per layer on CUDA-core MLPs) and the outD round trip.
#include "tc.cuh"
#include "params.cuh"
#include "kernels.h"

namespace abopt {

using namespace tc;

constexpr int OT_THREADS = 320, OT_ST = 4, OT_KCH = 8, OT_BK = 32;
constexpr int OT_A = 128 * OT_BK * 4;                     // 16 KB: 128 rows x 32 tf32
constexpr int OT_STAGE = 3 * OT_A;                        // 48 KB: A raw | B hi | B lo (the A operand is fed from tensor memory)
// TMEM columns of phase 1: main accumulator x 2 (promotion double buffer) | corrections (one, never promoted: its terms are 2^-11
// of the main ones, their truncation is irrelevant) | A ring: 2 slots x (feat hi 32 | lo 32)
constexpr uint32_t OT_TM_CORR = 256, OT_TM_A = 384;
// TMEM columns of phase 2: main accumulator | corrections | the layer's input activations hi (128) | lo (128) -- the three MLP
// layers run in TS mode too; the LayerNorm 1 output needed for the final residual is parked in (idle) shared memory instead
constexpr uint32_t OT_TM_ACT = 256;
constexpr int OT_ACT_KB = 2 * OT_A;                       // phase 2: one activation k-block, hi | lo
constexpr int OT_W_OFF = 4 * OT_ACT_KB;                   // phase 2: weight stages start behind the 4 activation k-blocks
constexpr int OT_BAR_OFF = OT_ST * OT_STAGE;              // 192 KB
constexpr int OT_EXCH_OFF = OT_BAR_OFF + 256;             // LayerNorm partial sums: [2 kinds][2 halves][128 rows]
constexpr int OT_SMEM = OT_EXCH_OFF + 4 * 128 * 4 + 1024;

struct TailArgs {
  const float* x; const uint8_t* mask;
  const float* bout; const float* ln1_g; const float* ln1_b;
  const float* b1; const float* b2; const float* b3;
  const float* ln2_g; const float* ln2_b;
  float* x_out; float* x_lo_out;
  const int* count;               // optional (focus mode): count[0] = number of valid rows, on the device (<= M)
};

// phase timestamps of CTA 0 (SM clock), read back by abopt_debug_clocks() slots 10..15
__device__ long long g_tail_clk[6];
__device__ __forceinline__ void tstamp(int slot) { if (blockIdx.x == 0) g_tail_clk[slot] = clock64(); }
void tail_debug_clocks(long long* out6) { cudaMemcpyFromSymbol(out6, g_tail_clk, sizeof(long long) * 6); }

__device__ __forceinline__ void epi_sync256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]; one thread issues ("TS" mode: the A operand is read from tensor memory, lane = row, column = k)
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// registers -> 32 lanes x 16 consecutive fp32 columns (thread = lane); the caller waits (tcgen05.wait::st) once for a batch
__device__ __forceinline__ void tmem_st16_nw(uint32_t taddr, const float (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                 "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                 "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
               : "memory");
}

__global__ void __launch_bounds__(OT_THREADS, 1)
outT_tail_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                 const __grid_constant__ CUtensorMap tmBl, const __grid_constant__ CUtensorMap tmWh,
                 const __grid_constant__ CUtensorMap tmWl, const __grid_constant__ CUtensorMap tmX,
                 const __grid_constant__ CUtensorMap tmXo, const __grid_constant__ CUtensorMap tmXl, int M, int K, const TailArgs ta) {
  extern __shared__ unsigned char smem_raw[];
  if (ta.count) {                              // focus mode: the row count lives on the device; surplus CTAs leave at once
    const int rows = ta.count[0];
    if ((int)blockIdx.x * 128 >= rows) return;
    M = rows < M ? rows : M;
  }
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + OT_BAR_OFF);
  uint64_t* empty = full + OT_ST;
  uint64_t* split = empty + OT_ST;
  uint64_t* tmem_full = split + OT_ST;         // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint64_t* w_full = tmem_empty + 2;           // [2]
  uint64_t* w_empty = w_full + 2;              // [2]
  uint64_t* act_ready = w_empty + 2;
  uint64_t* acc_full = act_ready + 1;
  uint64_t* x_full = acc_full + 1;
  uint64_t* ta_free = x_full + 1;              // [2]  the MMAs that read TMEM A slot s have completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ta_free + 2);
  float* exch_sum = reinterpret_cast<float*>(smem + OT_EXCH_OFF);      // [2][128]
  float* exch_sq = exch_sum + 256;                                     // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int nkb = K / OT_BK;
  const int nchunk = (nkb + OT_KCH - 1) / OT_KCH;
  constexpr uint32_t ACC_COLS = 256;                                   // main 128 | corrections 128

  if (threadIdx.x == 0) {
    for (int s = 0; s < OT_ST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&split[s], 8); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 8); mbar_init(&w_full[b], 1); mbar_init(&w_empty[b], 1); }
    mbar_init(act_ready, 8); mbar_init(acc_full, 1); mbar_init(x_full, 1);
    mbar_init(&ta_free[0], 1); mbar_init(&ta_free[1], 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmBh); tma_prefetch_desc(&tmBl); tma_prefetch_desc(&tmWh); tma_prefetch_desc(&tmWl);
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmXo); tma_prefetch_desc(&tmXl);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int last_c = nchunk - 1;
  if (threadIdx.x == 64) tstamp(0);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % OT_ST;
        // x tile of phase 2: pull it into L2 while the main loop runs
        if (kb == nkb / 2)
          for (int k2 = 0; k2 < 4; ++k2)
            asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(&tmX), "r"(k2 * OT_BK), "r"(m0) : "memory");
        mbar_wait(&empty[s], ((kb / OT_ST) & 1) ^ 1);
        unsigned char* st = smem + s * OT_STAGE;
        mbar_expect_tx(&full[s], 3 * OT_A);
        tma_load_2d(st, &tmA, kb * OT_BK, m0, &full[s]);
        tma_load_2d(st + OT_A, &tmBh, kb * OT_BK, 0, &full[s]);
        tma_load_2d(st + 2 * OT_A, &tmBl, kb * OT_BK, 0, &full[s]);
      }
      // phase 2: the pipeline memory is free once every out_transform MMA has retired
      mbar_wait(&tmem_full[last_c & 1], (last_c >> 1) & 1);
      // the block input x (residual) of this row tile lands in the hi slots of the four activation k-blocks, in the
      // swizzled layout the epilogue threads use anyway: no uncoalesced global loads
      mbar_expect_tx(x_full, 4 * OT_A);
      for (int kb = 0; kb < 4; ++kb) tma_load_2d(smem + kb * OT_ACT_KB, &tmX, kb * OT_BK, m0, x_full);
      for (int g = 0; g < 12; ++g) {
        const int l = g >> 2, kb = g & 3, s = g & 1;
        mbar_wait(&w_empty[s], ((g >> 1) & 1) ^ 1);
        unsigned char* st = smem + OT_W_OFF + s * OT_ACT_KB;
        mbar_expect_tx(&w_full[s], 2 * OT_A);
        tma_load_2d(st, &tmWh, kb * OT_BK, l * 128, &w_full[s]);
        tma_load_2d(st + OT_A, &tmWl, kb * OT_BK, l * 128, &w_full[s]);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = idesc_tf32(128, 128);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % OT_ST;
      const int c = kb / OT_KCH, buf = c & 1;
      const bool first = (kb % OT_KCH) == 0, last = (kb % OT_KCH) == OT_KCH - 1 || kb == nkb - 1;
      if (first && c >= 2) mbar_wait(&tmem_empty[buf], ((c >> 1) - 1) & 1);
      mbar_wait(&split[s], (kb / OT_ST) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t b_hi = smem_u32(smem + s * OT_STAGE) + OT_A, b_lo = b_hi + OT_A;
        const uint32_t d_main = tmem_base + buf * 128, d_small = tmem_base + OT_TM_CORR;
        const uint32_t ta = tmem_base + OT_TM_A + (kb & 1) * 64;
#pragma unroll
        for (int k = 0; k < OT_BK / 8; ++k) {
          const uint32_t ah = ta + k * 8, al = ah + 32;      // TS mode: A from tensor memory, column = k
          const uint64_t dbh = smem_desc_sw128(b_hi + k * 32), dbl = smem_desc_sw128(b_lo + k * 32);
          mma_tf32_ts(d_main, ah, dbh, idesc, (first && k == 0) ? 0u : 1u);
          mma_tf32_ts(d_small, ah, dbl, idesc, (kb == 0 && k == 0) ? 0u : 1u);
          mma_tf32_ts(d_small, al, dbh, idesc, 1u);
        }
        mma_commit(&ta_free[kb & 1]);
        mma_commit(&empty[s]);
        if (last) mma_commit(&tmem_full[buf]);
      }
      __syncwarp();
    }
    // phase 2: three 128 x 128 x 128 layers, A = activations written by the epilogue warps, accumulators in TMEM buffer 0
    for (int l = 0; l < 3; ++l) {
      mbar_wait(act_ready, l & 1);
      tc_fence_after();
      for (int kb = 0; kb < 4; ++kb) {
        const int g = l * 4 + kb, s = g & 1;
        mbar_wait(&w_full[s], (g >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_hi = smem_u32(smem + OT_W_OFF + s * OT_ACT_KB), b_lo = b_hi + OT_A;
          const uint32_t d_main = tmem_base, d_small = tmem_base + 128;
#pragma unroll
          for (int k = 0; k < OT_BK / 8; ++k) {
            const uint32_t ah = tmem_base + OT_TM_ACT + kb * OT_BK + k * 8, al = ah + 128;      // TS mode
            const uint64_t dbh = smem_desc_sw128(b_hi + k * 32), dbl = smem_desc_sw128(b_lo + k * 32);
            const uint32_t acc = (kb == 0 && k == 0) ? 0u : 1u;
            mma_tf32_ts(d_main, ah, dbh, idesc, acc);
            mma_tf32_ts(d_small, ah, dbl, idesc, acc);
            mma_tf32_ts(d_small, al, dbh, idesc, 1u);
          }
          mma_commit(&w_empty[s]);
          if (kb == 3) mma_commit(acc_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== splitters + epilogue (warps 2..9) =====================
    const int q = warp & 3;                                            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;                                  // which 64 of the 128 columns
    const int et = (warp - 2) * 32 + lane;                             // 0..255
    const int te = q * 32 + lane;                                      // row in the tile
    const int row = m0 + te;
    const int c0 = half * 64;
    float v[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = 0.f;
    auto promote = [&](int c) {
      const int buf = c & 1;
      mbar_wait(&tmem_full[buf], (c >> 1) & 1);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 128 + c0;
#pragma unroll
      for (int cc = 0; cc < 64; cc += 32) {
        float tm[32];
        tmem_ld_32x32(tbase + cc, tm);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[cc + i] += tm[i];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    };
    int promoted = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % OT_ST;
      mbar_wait(&full[s], (kb / OT_ST) & 1);
      // this thread's 16 floats of its feat row (k-block box [128 rows][32 k], 128-byte swizzle: 16-byte unit u sits at
      // u ^ (te & 7)) -> tf32 hi | lo in TMEM A slot kb & 1 (lane = row, column = k)
      const unsigned char* ar = smem + s * OT_STAGE + te * 128;
      float hi[16], lo[16];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 x = *reinterpret_cast<const float4*>(ar + (((half * 4 + u) ^ (te & 7)) << 4));
        hi[4 * u] = x.x; hi[4 * u + 1] = x.y; hi[4 * u + 2] = x.z; hi[4 * u + 3] = x.w;
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) lo[i] = tf32_lo(hi[i]);
      mbar_wait(&ta_free[kb & 1], ((kb >> 1) & 1) ^ 1);    // the MMAs of k-block kb - 2 have read this slot
      tc_fence_after();
      const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + OT_TM_A + (kb & 1) * 64 + half * 16;
      tmem_st16_nw(ta, hi);                                // (raw fp32: the tensor core ignores the low 13 mantissa bits)
      tmem_st16_nw(ta + 32, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&split[s]);
      if (kb % OT_KCH == 2 && kb / OT_KCH - 1 == promoted && kb >= OT_KCH) { promote(promoted); ++promoted; }
    }
    while (promoted < nchunk) { promote(promoted); ++promoted; }
    {
      // the correction accumulator (hi*lo + lo*hi of all k-blocks), once: every MMA has completed (last tmem_full)
      const uint32_t tcorr = tmem_base + ((uint32_t)(q * 32) << 16) + OT_TM_CORR + c0;
#pragma unroll
      for (int cc = 0; cc < 64; cc += 32) {
        float ts[32];
        tmem_ld_32x32(tcorr + cc, ts);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[cc + i] += ts[i];
      }
    }
    if (et == 0) tstamp(1);

    // ---------------- phase 2 ----------------
    // LayerNorm over the 128 columns of a row held by two threads (common/layers.py:146-155: biased variance, eps inside
    // the sqrt): partial sums exchanged through shared memory
    auto layer_norm = [&](float (&h)[64], const float* gamma, const float* beta) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 64; ++i) s += h[i];
      exch_sum[half * 128 + te] = s;
      epi_sync256();
      const float mean = (exch_sum[te] + exch_sum[128 + te]) * (1.f / 128.f);
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < 64; ++i) { h[i] -= mean; sq += h[i] * h[i]; }
      exch_sq[half * 128 + te] = sq;
      epi_sync256();
      const float inv = 1.f / sqrtf((exch_sq[te] + exch_sq[128 + te]) * (1.f / 128.f) + 1e-10f);
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c0 + i)), b = __ldg(reinterpret_cast<const float4*>(beta + c0 + i));
        h[i] = h[i] * inv * g.x + b.x; h[i + 1] = h[i + 1] * inv * g.y + b.y;
        h[i + 2] = h[i + 2] * inv * g.z + b.z; h[i + 3] = h[i + 3] * inv * g.w + b.w;
      }
    };
    // this thread's 64 values <-> rows te of the k-blocks 2 half, 2 half + 1: a row of a k-block is 128 bytes, its 16-byte
    // chunk c sits at (c ^ (te & 7)) (128-byte swizzle: what TMA writes / reads and what the UMMA descriptor expects)
    auto store_planes = [&](const float (&h)[64]) {
#pragma unroll
      for (int kbl = 0; kbl < 2; ++kbl) {
        unsigned char* base = smem + (2 * half + kbl) * OT_ACT_KB + te * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int i = kbl * 32 + c * 4;
          const uint32_t off = (uint32_t)((c ^ (te & 7)) << 4);
          *reinterpret_cast<float4*>(base + off) = make_float4(h[i], h[i + 1], h[i + 2], h[i + 3]);
          *reinterpret_cast<float4*>(base + OT_A + off) = make_float4(tf32_lo(h[i]), tf32_lo(h[i + 1]), tf32_lo(h[i + 2]), tf32_lo(h[i + 3]));
        }
      }
      fence_async_smem();                                              // generic-proxy writes -> visible to the async proxy
    };
    // the input of an MLP layer: this thread's 64 values as tf32 hi | lo into the TMEM activation columns (lane = row)
    auto store_act = [&](const float (&h)[64]) {
      const uint32_t ta_ = tmem_base + ((uint32_t)(q * 32) << 16) + OT_TM_ACT + c0;
#pragma unroll
      for (int cc = 0; cc < 64; cc += 16) {
        float hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { hi[i] = h[cc + i]; lo[i] = tf32_lo(h[cc + i]); }
        tmem_st16_nw(ta_ + cc, hi);
        tmem_st16_nw(ta_ + 128 + cc, lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(act_ready);
    };
    float* park = reinterpret_cast<float*>(smem);                      // [64][256]: value i of thread et (idle pipeline memory)
    const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + c0;
    // out_transform bias, mask_zero (layers.py:6-7), residual (x from the TMA-loaded tile), LayerNorm 1
    {
      const bool mk = row < M && ta.mask[row] != 0;
      mbar_wait(x_full, 0);
#pragma unroll
      for (int kbl = 0; kbl < 2; ++kbl) {
        const unsigned char* base = smem + (2 * half + kbl) * OT_ACT_KB + te * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int i = kbl * 32 + c * 4;
          const float4 xv = *reinterpret_cast<const float4*>(base + ((c ^ (te & 7)) << 4));
          const float4 bo = __ldg(reinterpret_cast<const float4*>(ta.bout + c0 + i));
          v[i] = xv.x + (mk ? v[i] + bo.x : 0.f); v[i + 1] = xv.y + (mk ? v[i + 1] + bo.y : 0.f);
          v[i + 2] = xv.z + (mk ? v[i + 2] + bo.z : 0.f); v[i + 3] = xv.w + (mk ? v[i + 3] + bo.w : 0.f);
        }
      }
    }
    layer_norm(v, ta.ln1_g, ta.ln1_b);
    // the LayerNorm 1 output is needed again for the residual at the end: park it in shared memory (every thread has read its
    // row of the x tile before the barriers inside layer_norm, so the region is free)
#pragma unroll
    for (int i = 0; i < 64; ++i) park[i * 256 + et] = v[i];
    store_act(v);
    if (et == 0) tstamp(2);
    for (int l = 0; l < 3; ++l) {
      mbar_wait(acc_full, l & 1);
      tc_fence_after();
      const float* bias = l == 0 ? ta.b1 : (l == 1 ? ta.b2 : ta.b3);
#pragma unroll
      for (int cc = 0; cc < 64; cc += 32) {
        float tm[32], ts[32];
        tmem_ld_32x32(tbase + cc, tm);
        tmem_ld_32x32(tbase + 128 + cc, ts);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c0 + cc + i));
          v[cc + i] = (tm[i] + ts[i]) + b.x; v[cc + i + 1] = (tm[i + 1] + ts[i + 1]) + b.y;
          v[cc + i + 2] = (tm[i + 2] + ts[i + 2]) + b.z; v[cc + i + 3] = (tm[i + 3] + ts[i + 3]) + b.w;
        }
      }
      tc_fence_before();
      if (l < 2) {
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] = fmaxf(v[i], 0.f);
        store_act(v);
      }
    }
    if (et == 0) tstamp(3);
    // residual (LayerNorm 1 output, parked in shared memory) + LayerNorm 2 (its barriers separate these reads from the output
    // planes written into the same region below)
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] += park[i * 256 + et];
    tc_fence_before();
    layer_norm(v, ta.ln2_g, ta.ln2_b);
    // x_out and its tf32 lo plane leave through the (drained) activation slots and TMA tensor stores: whole 128-byte
    // lines, rows >= M clipped by the hardware
    store_planes(v);
    epi_sync256();
    if (et == 0) {
      for (int kb = 0; kb < 4; ++kb) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(&tmXo), "r"(smem_u32(smem + kb * OT_ACT_KB)), "r"(kb * OT_BK), "r"(m0) : "memory");
        if (ta.x_lo_out)
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(&tmXl), "r"(smem_u32(smem + kb * OT_ACT_KB + OT_A)), "r"(kb * OT_BK), "r"(m0) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");    // smem must outlive the reads
    }
  }
  __syncthreads();
  if (threadIdx.x == 64) tstamp(4);
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

cudaError_t tail_tc_init() {
  return cudaFuncSetAttribute(outT_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OT_SMEM);
}

// x_out = GABlock tail(feat, x); feat (M, 1824) raw fp32, weights as hi / lo planes (Wmlp = [W1; W2; W3], each [128][128])
bool launch_outT_tail(int M, const float* feat, const float* x, const uint8_t* mask, const BlockW& w, float* x_out,
                      float* x_lo_out, cudaStream_t st, const int* count) {
  CUtensorMap a, bh, bl, wh, wl, tx, txo, txl;
  if (!make_tmap(&tx, x, M, F, F, 128) || !make_tmap(&txo, x_out, M, F, F, 128) ||
      !make_tmap(&txl, x_lo_out ? x_lo_out : x_out, M, F, F, 128) ||
      !make_tmap(&a, feat, M, NFEAT, NFEAT, 128) || !make_tmap(&bh, w.Wout, F, NFEAT, NFEAT, 128) ||
      !make_tmap(&bl, w.Wout_lo, F, NFEAT, NFEAT, 128) || !make_tmap(&wh, w.Wmlp, 3 * F, F, F, 128) ||
      !make_tmap(&wl, w.Wmlp_lo, 3 * F, F, F, 128))
    return false;
  ProfScope prof__(KK_TAIL, st);
  const TailArgs ta{x, mask, w.bout, w.ln1_g, w.ln1_b, w.b1, w.b2, w.b3, w.ln2_g, w.ln2_b, x_out, x_lo_out, count};
  outT_tail_kernel<<<(M + 127) / 128, OT_THREADS, OT_SMEM, st>>>(a, bh, bl, wh, wl, tx, txo, txl, M, NFEAT, ta);
  return true;
}

}  // namespace abopt
