// attn_logits_tc_kernel: final attention logits and the softmax over key residues on the 5th-gen tensor cores.
//
// For one (complex b, head h, tile of 128 query residues) the CTA computes
//   D[i][j] = QA[i] . KB[j]            64-wide contraction: q.k / sqrt(32)  and  -2 c_h qp_i . kp_j   (3xTF32, TMEM accumulator)
//   l[i][j] = ((D + pair_bias(i, j)) + (rq[i] + rk[j])) * sqrt(1/3) - 1e5 [key j masked]              ga.py:81-112,166,23
//   alpha[i][:] = softmax_j l[i][:]                                                                       ga.py:24
// QA / KB / rq / rk come packed from the projection GEMM epilogue (k_tc.cu: EpiProjPack); pair_bias is the hoisted
// z . W_b (k_pair.cu: pair_bias_kernel).
//
// Warp roles (320 threads):
//   warp 0     TMA producer: the 128 x 64 query operand once, then key blocks of 128 residues through a 2-stage ring
//   warp 1     TMEM allocator + MMA issuer: per key block 8 k-steps x 3 products (hi*hi, hi*lo, lo*hi) of
//              tcgen05.mma.kind::tf32 M=128 N=128 K=8 into TMEM columns [128 nb, 128 nb + 128)
//   warps 2-9  epilogue: two threads per query row (TMEM lane), each owning half of the keys.  Three passes over the
//              row in TMEM, 32 columns at a time: (1) logits -> running max, written back with tcgen05.st,
//              (2) exp -> running sum, written back, (3) normalise and store alpha.  The only cross-thread traffic is
//              the exchange of the two half-row maxima / sums through shared memory.
#include "tc.cuh"
#include "params.cuh"
#include "kernels.h"

namespace abopt {

using namespace tc;

constexpr int AL_THREADS = 320;                        // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int AL_EPI = 256;                            // epilogue threads: two per query row (each takes half of the keys)
constexpr int AL_BM = 128, AL_BN = 128, AL_K = 64;
constexpr int AL_BOX_BYTES = 128 * 32 * 4;             // one TMA box: 128 rows x 32 floats = 16 KB
constexpr int AL_OPER_BYTES = 2 * AL_BOX_BYTES;        // 128 rows x 64 floats (two boxes along K)
constexpr int AL_MAXCOLS = 512;                        // TMEM columns = longest key axis
constexpr int AL_A_OFF = 0;                            // A hi | A lo
constexpr int AL_B_OFF = 2 * AL_OPER_BYTES;            // 2 stages x (B hi | B lo)
constexpr int AL_TAB_OFF = AL_B_OFF + 2 * 2 * AL_OPER_BYTES;   // rk[512] | pen[512] | row max [2][128] | row sum [2][128]
constexpr int AL_BAR_OFF = AL_TAB_OFF + 2 * AL_MAXCOLS * 4 + 4 * 128 * 4;
constexpr int AL_SMEM = AL_BAR_OFF + 128 + 1024;

struct AttnLogitsArgs {
  int L, Lp, b0;
  const float* rq; const float* rk;      // [N][H][L]
  const float* bias;                     // [N][H][L keys][Lp] (query index contiguous)
  const uint8_t* mask;                   // [N][L]
  float* alpha;                          // [chunk][H][L][Lp]
};

// phase timestamps of CTA (0,0,0) (SM clock), read back by abopt_debug_clocks(): a poor man's timeline
__device__ long long g_attn_clk[16];
__device__ __forceinline__ void stamp(int slot) {
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_attn_clk[slot] = clock64();
}
void attn_debug_clocks(long long* out16) { cudaMemcpyFromSymbol(out16, g_attn_clk, sizeof(long long) * 16); }

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(AL_EPI) : "memory"); }

// DUAL: the hi*hi products and the two correction products accumulate in SEPARATE TMEM accumulators (columns
// [0, ncols) and [ncols, 2 ncols)): the tensor core truncates the fp32 accumulator on every accumulation, so keeping the
// large term to 8 accumulations instead of 24 cuts that bias 3x.  Needs 2 ncols <= 512 TMEM columns, i.e. L <= 256.
template <bool DUAL>
__global__ void __launch_bounds__(AL_THREADS, 1)
attn_logits_tc_kernel(const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
                      const __grid_constant__ CUtensorMap tmKh, const __grid_constant__ CUtensorMap tmKl,
                      const __grid_constant__ CUtensorMap tmBias, const AttnLogitsArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* ck = reinterpret_cast<float*>(smem + AL_TAB_OFF);
  float* pen = ck + AL_MAXCOLS;
  float* xmax = pen + AL_MAXCOLS;         // [2][128]
  float* xsum = xmax + 2 * 128;           // [2][128]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + AL_BAR_OFF);
  uint64_t* b_full = a_full + 1;          // [2]
  uint64_t* b_empty = b_full + 2;         // [2]
  uint64_t* tmem_full = b_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = a.L, Lp = a.Lp;
  const int i0 = blockIdx.x * AL_BM, h = blockIdx.y, bl = blockIdx.z, b = a.b0 + bl;
  const int nblk = (L + AL_BN - 1) / AL_BN;
  const int ncols = nblk * AL_BN;
  const int need = DUAL ? 2 * ncols : ncols;
  const uint32_t tmem_cols = need <= 128 ? 128u : (need <= 256 ? 256u : 512u);
  const int row_base = (b * H + h) * L;                 // first row of this (b, h) in the [N*H*L][*] views

  if (threadIdx.x == 0) {
    mbar_init(a_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmQh); tma_prefetch_desc(&tmQl); tma_prefetch_desc(&tmKh); tma_prefetch_desc(&tmKl);
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) stamp(0);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      unsigned char* A = smem + AL_A_OFF;
      mbar_expect_tx(a_full, 2 * AL_OPER_BYTES);
      tma_load_2d(A, &tmQh, 0, row_base + i0, a_full);
      tma_load_2d(A + AL_BOX_BYTES, &tmQh, 32, row_base + i0, a_full);
      tma_load_2d(A + AL_OPER_BYTES, &tmQl, 0, row_base + i0, a_full);
      tma_load_2d(A + AL_OPER_BYTES + AL_BOX_BYTES, &tmQl, 32, row_base + i0, a_full);
      for (int nb = 0; nb < nblk; ++nb) {
        const int s = nb & 1;
        mbar_wait(&b_empty[s], ((nb >> 1) & 1) ^ 1);
        unsigned char* B = smem + AL_B_OFF + s * 2 * AL_OPER_BYTES;
        mbar_expect_tx(&b_full[s], 2 * AL_OPER_BYTES);
        const int r0 = row_base + nb * AL_BN;
        tma_load_2d(B, &tmKh, 0, r0, &b_full[s]);
        tma_load_2d(B + AL_BOX_BYTES, &tmKh, 32, r0, &b_full[s]);
        tma_load_2d(B + AL_OPER_BYTES, &tmKl, 0, r0, &b_full[s]);
        tma_load_2d(B + AL_OPER_BYTES + AL_BOX_BYTES, &tmKl, 32, r0, &b_full[s]);
        if (nb == (nblk > 1 ? 1 : 0)) {
          // once the first operand loads are queued: pull this tile of the (HBM-resident) pair bias into L2 while the
          // MMAs run -- boxes of [<= 256 keys][128 queries] of the [N*H*L keys][Lp queries] view
          for (int j0 = 0; j0 < L; j0 += 256)
            asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(&tmBias), "r"(i0), "r"(row_base + j0) : "memory");
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = idesc_tf32(AL_BM, AL_BN);
    mbar_wait(a_full, 0);
    if (lane == 0) stamp(1);
    for (int nb = 0; nb < nblk; ++nb) {
      const int s = nb & 1;
      mbar_wait(&b_full[s], (nb >> 1) & 1);
      if (lane == 0) stamp(2 + (nb & 1));
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_hi = smem_u32(smem + AL_A_OFF), a_lo = a_hi + AL_OPER_BYTES;
        const uint32_t b_hi = smem_u32(smem + AL_B_OFF + s * 2 * AL_OPER_BYTES), b_lo = b_hi + AL_OPER_BYTES;
        const uint32_t d = tmem_base + nb * AL_BN;
        const uint32_t dc = DUAL ? d + ncols : d;          // correction accumulator
#pragma unroll
        for (int kk = 0; kk < AL_K / 8; ++kk) {          // UMMA_K = 8 tf32 = 32 B inside a 128 B swizzle row; 4 steps per box
          const uint32_t ko = (kk >> 2) * AL_BOX_BYTES + (kk & 3) * 32;
          const uint64_t dah = smem_desc_sw128(a_hi + ko), dal = smem_desc_sw128(a_lo + ko);
          const uint64_t dbh = smem_desc_sw128(b_hi + ko), dbl = smem_desc_sw128(b_lo + ko);
          mma_tf32(d, dah, dbh, idesc, kk == 0 ? 0u : 1u);
          mma_tf32(dc, dah, dbl, idesc, (DUAL && kk == 0) ? 0u : 1u);
          mma_tf32(dc, dal, dbh, idesc, 1u);
        }
        mma_commit(&b_empty[s]);
        if (nb == nblk - 1) mma_commit(tmem_full);
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;                               // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;                     // which half of the key axis this thread handles
    const int te = q * 32 + lane;                         // 0..127: query row in the tile
    const int i = i0 + te;
    const bool valid = i < L;
    const int cbeg = half * (ncols / 2), cend = cbeg + ncols / 2;
    // per-key tables: rk[j] and the mask penalty (1e5 for masked keys, +inf beyond the end -> alpha = 0)
    for (int j = te + half * 128; j < ncols; j += AL_EPI) {
      ck[j] = (j < L) ? __ldg(a.rk + (size_t)row_base + j) : 0.f;
      pen[j] = (j < L) ? (a.mask[(size_t)b * L + j] != 0 ? 0.f : 1e5f) : INFINITY;
    }
    const float rqi = valid ? __ldg(a.rq + (size_t)row_base + i) : 0.f;
    const float* bias_col = a.bias + (size_t)row_base * Lp + (valid ? i : 0);     // + j * Lp: bias is stored [j][i]
    float* alpha_row = a.alpha + ((size_t)(bl * H + h) * L + (valid ? i : 0)) * Lp;
    float bv[32], bn[32];
    auto load_bias = [&](int c0, float (&dst)[32]) {
#pragma unroll
      for (int e = 0; e < 32; ++e) dst[e] = (c0 + e < L) ? __ldg(bias_col + (size_t)(c0 + e) * Lp) : 0.f;
    };
    epi_sync();                                           // tables visible to all epilogue warps
    if (te == 0 && half == 0) stamp(4);
    mbar_wait(tmem_full, 0);
    if (te == 0 && half == 0) stamp(5);
    tc_fence_after();
    load_bias(cbeg, bv);                                  // by now the tile's bias has been prefetched into L2
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const float scale = 0.57735026918962576f;             // sqrt(1/3), ga.py:166
    const float l2e = 1.4426950408889634f;

    // ---- pass 1: logits, running max.  The pair bias of the next 32 keys is in flight while this chunk is processed.
    float m = -INFINITY;
    for (int c0 = cbeg; c0 < cend; c0 += 32) {
      if (c0 + 32 < cend) load_bias(c0 + 32, bn);
      float v[32];
      tmem_ld_32x32(trow + c0, v);
      if (DUAL) {
        float w[32];
        tmem_ld_32x32(trow + ncols + c0, w);
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] += w[e];
      }
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const float lgt = ((v[e] + bv[e]) + (rqi + ck[c0 + e])) * scale - pen[c0 + e];
        m = fmaxf(m, lgt);
        v[e] = lgt;
      }
      tmem_st_32x32(trow + c0, v);
#pragma unroll
      for (int e = 0; e < 32; ++e) bv[e] = bn[e];
    }
    xmax[half * 128 + te] = m;
    if (te == 0 && half == 0) stamp(6);
    epi_sync();
    m = fmaxf(xmax[te], xmax[128 + te]);
    // ---- pass 2: exp(l - m) = 2^(l log2e - m log2e), running sum
    const float ml2e = m * l2e;
    float sum = 0.f;
    for (int c0 = cbeg; c0 < cend; c0 += 32) {
      float v[32];
      tmem_ld_32x32(trow + c0, v);
#pragma unroll
      for (int e = 0; e < 32; ++e) { v[e] = exp2f(fmaf(v[e], l2e, -ml2e)); sum += v[e]; }
      tmem_st_32x32(trow + c0, v);
    }
    xsum[half * 128 + te] = sum;
    if (te == 0 && half == 0) stamp(7);
    epi_sync();
    sum = xsum[te] + xsum[128 + te];
    // ---- pass 3: normalise, store
    const float inv = 1.0f / sum;
    for (int c0 = cbeg; c0 < cend; c0 += 32) {
      float v[32];
      tmem_ld_32x32(trow + c0, v);
      if (valid) {
        // 256-bit stores (rows are 32-byte aligned, Lp % 8 == 0): every store fills a whole 32 B sector
#pragma unroll
        for (int e = 0; e < 32; e += 8)
          if (c0 + e < Lp)
            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(alpha_row + c0 + e), "f"(v[e] * inv),
                         "f"(v[e + 1] * inv), "f"(v[e + 2] * inv), "f"(v[e + 3] * inv), "f"(v[e + 4] * inv), "f"(v[e + 5] * inv),
                         "f"(v[e + 6] * inv), "f"(v[e + 7] * inv) : "memory");
      }
    }
    if (te == 0 && half == 0) stamp(8);
    tc_fence_before();
  }
  __syncthreads();
  if (threadIdx.x == 0) stamp(9);
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

cudaError_t attn_tc_init() {
  cudaError_t e = cudaFuncSetAttribute(attn_logits_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AL_SMEM);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(attn_logits_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AL_SMEM);
}

bool launch_attn_logits_tc(int nb, int b0, int N, int L, int Lp, const AttnOperands& op, const float* bias_layer, const uint8_t* mask,
                           float* alpha, cudaStream_t st) {
  if (L > AL_MAXCOLS) return false;
  CUtensorMap qh, ql, kh, kl, bm;
  const uint64_t rows = (uint64_t)N * H * L;
  if (!make_tmap(&qh, op.QA, rows, 64, 64, 128) || !make_tmap(&ql, op.QA_lo, rows, 64, 64, 128) ||
      !make_tmap(&kh, op.KB, rows, 64, 64, 128) || !make_tmap(&kl, op.KB_lo, rows, 64, 64, 128))
    return false;
  // pair bias as a plain (unswizzled) 2-D tensor for L2 prefetches: [N*H*L keys][Lp queries], box [<=256][<=128]
  if (!make_tmap_plain(&bm, bias_layer, rows, (uint64_t)Lp, (uint64_t)Lp, rows < 256 ? (uint32_t)rows : 256u, Lp < 128 ? (uint32_t)Lp : 128u))
    return false;
  ProfScope prof__(KK_LOGITS, st);
  AttnLogitsArgs a{L, Lp, b0, op.rq, op.rk, bias_layer, mask, alpha};
  dim3 grid((L + AL_BM - 1) / AL_BM, H, nb);
  const int ncols = ((L + AL_BN - 1) / AL_BN) * AL_BN;
  if (2 * ncols <= AL_MAXCOLS) attn_logits_tc_kernel<true><<<grid, AL_THREADS, AL_SMEM, st>>>(qh, ql, kh, kl, bm, a);
  else attn_logits_tc_kernel<false><<<grid, AL_THREADS, AL_SMEM, st>>>(qh, ql, kh, kl, bm, a);
  return true;
}

}  // namespace abopt
