// Pair featurisation (the step before the sampling loop): PairEmbedding.forward,
// /root/reference/AbDock/src/modules/encoders/pair.py:37-101 (AbDesign: diffab/modules/encoders/pair.py, same lines), with
// pairwise_dihedrals (modules/common/geometry.py:351-376), dihedral_from_four_points (:254-271) and AngularEncoding
// (modules/common/layers.py:85-106), as ONE persistent kernel that writes pair_feat (N,L,L,64) once and never materialises the
// (N,L,L,A*A) distance tensor (3.8 GB at N=64, L=256, A=15), the (N,L,L,218) concatenation or any MLP intermediate.
//
// Work unit = one query residue (n, i): its atoms, its 22 rows of the softplus'd distance-coefficient table and its scalars are
// staged once, then the L keys are walked in tiles of 128 pairs.  Per tile, on chip only:
//   g[ab][pair] = mask_a mask_b exp(-softplus(coef[aa_i aa_j][ab]) (|x_ia - x_jb| / 10)^2)          pair.py:77-84
//   h1 = relu(Wd1 g + b), h2 = relu(Wd2 h1 + b) * structure_pair                                      pair.py:84-87
//   phi/psi -> [x, sin(x f), cos(x f)] * structure_pair                                               pair.py:90-94
//   o1 = relu(T_aa[aa_i aa_j] + same_chain T_rel[clamp(res_i - res_j)] + W1[:,128:192] h2 + W1[:,192:218] ang + b1)
//        (the aa-pair and relative-position embeddings go through the first out_mlp layer as pre-multiplied tables)  pair.py:65-74,97-98
//   o2 = relu(W2 o1 + b2), z = (W3 o2 + b3) * has_CA_i has_CA_j                                        pair.py:98-99
// The five dense layers run on the 5th-gen tensor cores as 3xTF32 (a b + a_lo b + a b_lo) because the parity target is the
// reference's fp32 output; see pair_embed_tc_kernel below.  History: FP32 FFMA 18.4 ms -> mma.sync m16n8k8 tf32
// 10.2 ms (round 1) -> tcgen05 TS mode 4.1 ms at B=64, L=256, 15 atoms.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/abopt_b200.h"
#include "kernels.h"
#include "tc.cuh"

namespace abopt {
int api_fail(int code, const std::string& msg);      // api.cu: sets abopt_last_error()

namespace {
constexpr int PE_MAXA = 15;          // max_num_heavyatoms (utils/protein/constants.py:143)
constexpr int PE_AA = 22;            // max_aa_types (pair.py:12)
constexpr int PE_RELPOS = 32;        // max_relpos (pair.py:12)
constexpr int PE_UNK = 20;           // AA.UNK (constants.py:108)
constexpr int PE_ANG = 26;           // AngularEncoding.get_out_dim(2) (layers.py:94-95)

struct PairEmbedW {                  // device pointers into one packed allocation
  int A, A2;
  const float* Taa;                  // [484][64]   aa_pair_embed . W1[:, 0:64]^T
  const float* Trel;                 // [65][64]    relpos_embed  . W1[:, 64:128]^T
  const float* bias;                 // [5][64]     bd1, bd2, b1, b2, b3
  float freq[6];                     // dihedral_embed.freq_bands
  // B-operand boxes [64 out][32 k] (128-byte rows, 16-byte units XOR-swizzled by out & 7), hi plane | lo plane,
  // 16 KB each: kb1 boxes of distance_embed.0, then 9 resident ones (distance_embed.2: 2, out_mlp.0 h2 part: 2, angle part: 1
  // with the two angles' 13 features at k 0..12 and 16..28, out_mlp.2: 2, out_mlp.4: 2)
  const float* boxes;
  int kb1;
  const float* coefT;                // [484][A][16]  softplus(aapair_to_distcoef) as [aa pair][key atom][query atom, padded to 16]
};

struct PairEmbedArgs {
  int N, L, A_in;
  const long long* aa; const long long* res_nb; const long long* chain_nb;
  const float* pos; const uint8_t* mask_atoms; const uint8_t* structure_mask; const uint8_t* sequence_mask;
  float* out;
};

// ------------------------------------------------------------------------------------------ the kernel
// pair_embed_tc_kernel.  Tile = one query residue x 128 keys; the five dense
// layers are 128 x 64 x K GEMMs as 3xTF32 tcgen05.mma in TS mode: the A operand (activations, tf32 hi | lo) lives in TENSOR MEMORY
// and is written by the threads that produce it (thread = pair x half of the columns: tcgen05.st), the B operand (weights) comes
// from shared memory as pre-swizzled K-major boxes [64 out][32 k] hi | lo packed at finalize(): distance_embed.2 and the three
// out_mlp layers are resident (9 boxes, 144 KB), distance_embed.0 (K = A^2, 8 boxes for 15 atoms) streams from L2 through a 2-stage
// bulk-copy ring.  10 warps: TMA producer, MMA issuer, 8 compute warps; accumulators (main | corrections) in TMEM.
//   layer 1  K order = [key atom b][query atom a, padded to 16]: a k-block of 32 entries is two key atoms, a thread evaluates the
//            16 Gaussians of ONE key atom of its pair against the query atoms it keeps in registers -> TMEM A ring (2 slots) -> 12 MMAs
//   layers 2-5  acc -> bias / ReLU / masks (+ tables, angular features) in registers -> TMEM activations -> 24-36 MMAs -> acc
constexpr int PT_THREADS = 320, PT_TILE = 128, PT_NST = 2, PT_NRES = 9;
constexpr int PT_BOX = 64 * 32 * 4, PT_BOX2 = 2 * PT_BOX;      // one plane of a weight box (8 KB); hi | lo
constexpr int PT_RES_OFF = 0, PT_RING_OFF = PT_NRES * PT_BOX2, PT_COEF_OFF = PT_RING_OFF + PT_NST * PT_BOX2;
constexpr int PT_POSJ_OFF = PT_COEF_OFF + PE_AA * PE_MAXA * 16 * 4, PT_POSI_OFF = PT_POSJ_OFF + PT_TILE * PE_MAXA * 3 * 4;
constexpr int PT_INT_OFF = PT_POSI_OFF + 192, PT_BIAS_OFF = PT_INT_OFF + 5 * PT_TILE * 4 + 64, PT_BAR_OFF = PT_BIAS_OFF + 5 * 64 * 4;
constexpr int PT_SMEM = PT_BAR_OFF + 160 + 1024;
static_assert(PT_SMEM <= 227 * 1024, "pair_embed_tc_kernel: shared memory");
constexpr uint32_t PT_TM_A1 = 0, PT_TM_ACC = 128, PT_TM_ACT = 256, PT_TM_ACTLO = 352;   // TMEM columns

__device__ __forceinline__ void pt_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void pt_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void pt_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                 "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                 "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
               : "memory");
}
// hi | lo planes of 16 values -> TMEM columns col.. of the activation (or A-ring) region; lo_off = distance of the lo plane
__device__ __forceinline__ void pt_store16(uint32_t t_hi, uint32_t lo_off, const float (&v)[16]) {
  float lo[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) lo[i] = tf32_lo(v[i]);
  pt_st16(t_hi, v);
  pt_st16(t_hi + lo_off, lo);
}
// main + corrections, 16 columns
__device__ __forceinline__ void pt_ld16_sum(uint32_t t_main, float (&v)[16]) {
  uint32_t m[16], c[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(m[0]), "=r"(m[1]), "=r"(m[2]), "=r"(m[3]), "=r"(m[4]), "=r"(m[5]), "=r"(m[6]), "=r"(m[7]), "=r"(m[8]),
                 "=r"(m[9]), "=r"(m[10]), "=r"(m[11]), "=r"(m[12]), "=r"(m[13]), "=r"(m[14]), "=r"(m[15]) : "r"(t_main));
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7]), "=r"(c[8]),
                 "=r"(c[9]), "=r"(c[10]), "=r"(c[11]), "=r"(c[12]), "=r"(c[13]), "=r"(c[14]), "=r"(c[15]) : "r"(t_main + 64));
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(m[0]), "+r"(m[1]), "+r"(m[2]), "+r"(m[3]), "+r"(m[4]), "+r"(m[5]), "+r"(m[6]), "+r"(m[7]), "+r"(m[8]),
                 "+r"(m[9]), "+r"(m[10]), "+r"(m[11]), "+r"(m[12]), "+r"(m[13]), "+r"(m[14]), "+r"(m[15]) :: "memory");
  asm volatile("" : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7]), "+r"(c[8]),
                    "+r"(c[9]), "+r"(c[10]), "+r"(c[11]), "+r"(c[12]), "+r"(c[13]), "+r"(c[14]), "+r"(c[15]) :: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(m[i]) + __uint_as_float(c[i]);
}

__global__ void __launch_bounds__(PT_THREADS, 1) pair_embed_tc_kernel(PairEmbedW w, PairEmbedArgs a) {
  using namespace tc;
  extern __shared__ unsigned char pt_smem_raw[];
  unsigned char* smem = pt_smem_raw + ((1024u - (smem_u32(pt_smem_raw) & 1023u)) & 1023u);
  float* sCoef = reinterpret_cast<float*>(smem + PT_COEF_OFF);      // [22][A][16]
  float* sPosJ = reinterpret_cast<float*>(smem + PT_POSJ_OFF);      // [128][A*3]
  float* sPosI = reinterpret_cast<float*>(smem + PT_POSI_OFF);      // [A*3] (48 slots)
  int* sAaJ = reinterpret_cast<int*>(smem + PT_INT_OFF);            // [128] amino-acid slot of the key
  int* sRel = sAaJ + PT_TILE;                                       // [128] row of T_rel, -1 = other chain
  int* sKeep = sRel + PT_TILE;                                      // [128] structure_mask_i & structure_mask_j
  int* sOk = sKeep + PT_TILE;                                       // [128] has_CA_i & has_CA_j (& j < L)
  int* sBitsJ = sOk + PT_TILE;                                      // [128] atom mask of the key, one bit per atom
  int* sMisc = sBitsJ + PT_TILE;                                    // [16] scalars of the query residue
  float* sBias = reinterpret_cast<float*>(smem + PT_BIAS_OFF);      // [5][64]
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + PT_BAR_OFF);      // [2]
  uint64_t* w_empty = w_full + 2;       // [2]
  uint64_t* a_ready = w_empty + 2;      // [2]  the compute warps have written A-ring slot s
  uint64_t* ta_free = a_ready + 2;      // [2]  the MMAs that read A-ring slot s have completed
  uint64_t* act_ready = ta_free + 2;    // the next layer's activations are in TMEM
  uint64_t* acc_full = act_ready + 1;   // a layer's MMAs have completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int A = w.A, A2 = w.A2, kb1 = w.kb1;
  const int L = a.L, A_in = a.A_in;
  const int n_tiles = (L + PT_TILE - 1) / PT_TILE;
  const long long rows = (long long)a.N * L;

  // ---- resident weight boxes (already in the UMMA layout) and biases: once per CTA
  {
    const float4* src = reinterpret_cast<const float4*>(w.boxes + (size_t)kb1 * (PT_BOX2 / 4));
    float4* dst = reinterpret_cast<float4*>(smem + PT_RES_OFF);
    for (int i = tid; i < PT_NRES * PT_BOX2 / 16; i += PT_THREADS) dst[i] = src[i];
    for (int i = tid; i < 5 * 64; i += PT_THREADS) sBias[i] = w.bias[i];
    fence_async_smem();
  }
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); mbar_init(&a_ready[s], 8); mbar_init(&ta_free[s], 1); }
    mbar_init(act_ready, 8); mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== producer: the boxes of distance_embed.0, one per k-block and tile =====================
    if (elect_one()) {
      int g = 0;
      for (long long row = blockIdx.x; row < rows; row += gridDim.x)
        for (int jt = 0; jt < n_tiles; ++jt)
          for (int kb = 0; kb < kb1; ++kb, ++g) {
            const int s = g & 1;
            mbar_wait(&w_empty[s], ((g >> 1) & 1) ^ 1);
            mbar_expect_tx(&w_full[s], PT_BOX2);
            asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(smem + PT_RING_OFF + s * PT_BOX2)), "l"(w.boxes + (size_t)kb * (PT_BOX2 / 4)), "r"(PT_BOX2),
                           "r"(smem_u32(&w_full[s])) : "memory");
          }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // (the lo plane of a weight box follows its hi plane, the corrections accumulator follows the main one: x_hi . W_hi and x_hi . W_lo
    // are one 128-wide instruction, ~77 clk instead of 2 x ~45)
    constexpr uint32_t idesc = idesc_tf32(128, 64), idesc128 = idesc_tf32(128, 128);
    const uint32_t d_main = tmem_base + PT_TM_ACC, d_corr = d_main + 64;
    int g = 0, ac = 0;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x)
      for (int jt = 0; jt < n_tiles; ++jt) {
        for (int kb = 0; kb < kb1; ++kb, ++g) {
          const int s = g & 1;
          mbar_wait(&a_ready[s], (g >> 1) & 1);
          mbar_wait(&w_full[s], (g >> 1) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t b_hi = smem_u32(smem + PT_RING_OFF + s * PT_BOX2);
            const uint32_t ta = tmem_base + PT_TM_A1 + s * 64;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t ah = ta + k * 8, al = ah + 32;
              const uint64_t dbh = smem_desc_sw128(b_hi + k * 32);
              const uint32_t acc = (kb == 0 && k == 0) ? 0u : 1u;
              pt_mma_ts(d_main, ah, dbh, idesc128, acc);        // a_hi . [W_hi | W_lo] -> [main | corrections], one 128-wide instruction
              pt_mma_ts(d_corr, al, dbh, idesc, 1u);
            }
            mma_commit(&ta_free[s]);
            mma_commit(&w_empty[s]);
            if (kb == kb1 - 1) mma_commit(acc_full);
          }
          __syncwarp();
        }
#pragma unroll 1
        for (int l = 0; l < 4; ++l, ++ac) {
          const int nkb = l == 1 ? 3 : 2, box0 = l == 0 ? 0 : (l == 1 ? 2 : (l == 2 ? 5 : 7));
          mbar_wait(act_ready, ac & 1);
          tc_fence_after();
          if (elect_one()) {
            for (int kb = 0; kb < nkb; ++kb) {
              const uint32_t b_hi = smem_u32(smem + PT_RES_OFF + (box0 + kb) * PT_BOX2);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t ah = tmem_base + PT_TM_ACT + kb * 32 + k * 8, al = tmem_base + PT_TM_ACTLO + kb * 32 + k * 8;
                const uint64_t dbh = smem_desc_sw128(b_hi + k * 32);
                const uint32_t acc = (kb == 0 && k == 0) ? 0u : 1u;
                pt_mma_ts(d_main, ah, dbh, idesc128, acc);
                pt_mma_ts(d_corr, al, dbh, idesc, 1u);
              }
            }
            mma_commit(acc_full);
          }
          __syncwarp();
        }
      }
  } else {
    // ===================== compute warps (2..9): thread = (pair p of the tile, half of the columns) =====================
    const int ct = tid - 64;                              // 0..255
    const int cw = warp - 2;                              // 0..7
    const int q = warp & 3;                               // TMEM lane quarter this warp may access
    const int p = q * 32 + lane;                          // pair of the tile = TMEM lane
    const int half = cw >> 2;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    int g = 0, af = 0;                                    // A-ring slot uses, acc_full completions consumed
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
      const int n = (int)(row / L);
      pt_bar();                                           // previous row fully consumed
      // ---- query residue
      if (ct < A * 3) sPosI[ct] = a.pos[((size_t)row * A_in) * 3 + ct];
      if (ct == 64) {
        long long aa = a.aa[row];
        if (a.sequence_mask && !a.sequence_mask[row]) aa = PE_UNK;                     // pair.py:62-64
        aa = aa < 0 ? 0 : (aa >= PE_AA ? PE_AA - 1 : aa);
        int bits = 0;
        for (int k = 0; k < A; ++k) bits |= (a.mask_atoms[(size_t)row * A_in + k] ? 1 : 0) << k;
        sMisc[0] = (int)aa;
        sMisc[1] = bits;
        sMisc[2] = a.structure_mask ? (a.structure_mask[row] ? 1 : 0) : 1;
      }
      pt_bar();
      const int aa_i = sMisc[0], bits_i = sMisc[1], keep_i = sMisc[2];
      const int ok_i = (bits_i >> 1) & 1;                                              // BBHeavyAtom.CA = 1, pair.py:57
      {
        const float* src = w.coefT + (size_t)aa_i * PE_AA * A * 16;
        const int total = PE_AA * A * 16;
        for (int k0 = ct; k0 < total; k0 += 4 * 256) {
          float val[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) val[u] = k0 + u * 256 < total ? __ldg(src + k0 + u * 256) : 0.f;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (k0 + u * 256 < total) sCoef[k0 + u * 256] = val[u];
        }
      }
      const long long res_i = a.res_nb[row], chain_i = a.chain_nb[row];

      for (int jt = 0; jt < n_tiles; ++jt) {
        const int j0 = jt * PT_TILE;
        pt_bar();                                         // previous tile's readers done (sCoef fill ordered too)
        // ---- stage the 128 keys: all 256 threads over the tile's coordinates, four independent loads in flight per thread
        {
          const int nf = A * 3, total = PT_TILE * nf;
          const float* src0 = a.pos + (((size_t)n * L + j0) * A_in) * 3;
          for (int e0 = ct; e0 < total; e0 += 4 * 256) {
            float val[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int e = e0 + u * 256;
              const int key = e / nf, c = e - key * nf;
              val[u] = (e < total && j0 + key < L) ? __ldg(src0 + (size_t)key * A_in * 3 + c) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (e0 + u * 256 < total) sPosJ[e0 + u * 256] = val[u];
          }
        }
        if (ct < PT_TILE) {
          const int j = j0 + ct;
          int aaj = 0, rel = -1, keep = 0, ok = 0, bits = 0;
          if (j < L) {
            const size_t rj = (size_t)n * L + j;
            long long aa = a.aa[rj];
            if (a.sequence_mask && !a.sequence_mask[rj]) aa = PE_UNK;
            aaj = (int)(aa < 0 ? 0 : (aa >= PE_AA ? PE_AA - 1 : aa));
            for (int k = 0; k < A; ++k) bits |= (a.mask_atoms[rj * A_in + k] ? 1 : 0) << k;
            long long d = res_i - a.res_nb[rj];                                        // pair.py:70-73
            d = d < -PE_RELPOS ? -PE_RELPOS : (d > PE_RELPOS ? PE_RELPOS : d);
            rel = (a.chain_nb[rj] == chain_i) ? (int)d + PE_RELPOS : -1;
            keep = keep_i && (a.structure_mask ? (a.structure_mask[rj] != 0) : 1);
            ok = ok_i && ((bits >> 1) & 1);
          }
          sAaJ[ct] = aaj; sRel[ct] = rel; sKeep[ct] = keep; sOk[ct] = ok; sBitsJ[ct] = bits;
        }
        pt_bar();
        const int aa_j = sAaJ[p], rel = sRel[p], bits_j = sBitsJ[p];
        const float keepf = sKeep[p] ? 1.f : 0.f;
        const float* xjb = sPosJ + p * (A * 3);

        // ---- layer 1: distance Gaussians -> TMEM A ring -> distance_embed.0.  k-block kb = key atoms 2 kb, 2 kb + 1; this thread:
        //      key atom ib = 2 kb + half of its pair against the 16 (padded) query atoms
        float xq[16][3];                                  // query atoms (registers; zero beyond A)
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const bool in = u < A;
          xq[u][0] = in ? sPosI[u * 3] : 0.f; xq[u][1] = in ? sPosI[u * 3 + 1] : 0.f; xq[u][2] = in ? sPosI[u * 3 + 2] : 0.f;
        }
        const float* cfp = sCoef + aa_j * (A * 16);
        for (int kb = 0; kb < kb1; ++kb, ++g) {
          const int ib = 2 * kb + half;
          const bool bon = ib < A && ((bits_j >> ib) & 1);
          const int ibc = ib < A ? ib : 0;
          const float xj0 = xjb[ibc * 3], xj1 = xjb[ibc * 3 + 1], xj2 = xjb[ibc * 3 + 2];
          const float4* c4 = reinterpret_cast<const float4*>(cfp + ibc * 16);
          float cf[16];
#pragma unroll
          for (int u = 0; u < 4; ++u) { const float4 c = c4[u]; cf[4 * u] = c.x; cf[4 * u + 1] = c.y; cf[4 * u + 2] = c.z; cf[4 * u + 3] = c.w; }
          float gv[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const float dx = xq[u][0] - xj0, dy = xq[u][1] - xj1, dz = xq[u][2] - xj2;
            const float d2 = (dx * dx + dy * dy + dz * dz) * 0.01f;                     // (|x_ia - x_jb| / 10)^2, pair.py:77-82
            const bool on = bon && u < A && ((bits_i >> u) & 1);
            const float v = expf(-cf[u] * d2);
            gv[u] = on ? v : 0.f;                                                         // pair.py:82-84
          }
          const int s = g & 1;
          mbar_wait(&ta_free[s], ((g >> 1) & 1) ^ 1);     // the MMAs of the k-block two before have read this slot
          tc_fence_after();
          pt_store16(tlane + PT_TM_A1 + s * 64 + half * 16, 32, gv);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_ready[s]);
        }
        // ---- inter-residue dihedral of this thread's angle (half 0: phi, half 1: psi) + angular encoding (13 values)
        float ang[16];
        {
          const float* Ni = sPosI; const float* CAi = sPosI + 3; const float* Ci = sPosI + 6;
          const float* Nj = xjb; const float* CAj = Nj + 3; const float* Cj = Nj + 6;
          const float x = half == 0 ? dihedral4(Ci, Nj, CAj, Cj) : dihedral4(Ni, CAi, Ci, Nj);   // geometry.py:362-373
          ang[0] = x * keepf;                                                                    // pair.py:92-94
#pragma unroll
          for (int f = 0; f < 6; ++f) {
            float sn, cs;
            sincosf(x * w.freq[f], &sn, &cs);
            ang[1 + f] = sn * keepf;
            ang[7 + f] = cs * keepf;
          }
          ang[13] = 0.f; ang[14] = 0.f; ang[15] = 0.f;
        }
        // table part of the first out_mlp layer (gathers in flight during the two layers before it)
        float tab[32];
        {
          const float4* ta4 = reinterpret_cast<const float4*>(w.Taa + ((size_t)(aa_i * PE_AA + aa_j)) * 64 + half * 32);
          const float4* tr4 = reinterpret_cast<const float4*>(w.Trel + (size_t)(rel < 0 ? 0 : rel) * 64 + half * 32);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 va = __ldg(ta4 + c);
            float4 vr = __ldg(tr4 + c);
            if (rel < 0) vr = make_float4(0.f, 0.f, 0.f, 0.f);
            tab[4 * c] = va.x + vr.x; tab[4 * c + 1] = va.y + vr.y; tab[4 * c + 2] = va.z + vr.z; tab[4 * c + 3] = va.w + vr.w;
          }
        }
        const uint32_t t_acc = tlane + PT_TM_ACC + half * 32, t_act = tlane + PT_TM_ACT + half * 32;
        // ---- layers 2-5: accumulator -> registers -> next activations in TMEM
#pragma unroll 1
        for (int l = 0; l < 5; ++l, ++af) {
          mbar_wait(acc_full, af & 1);
          tc_fence_after();
          float v[2][16];
          pt_ld16_sum(t_acc, v[0]);
          pt_ld16_sum(t_acc + 16, v[1]);
          const float* b = sBias + l * 64 + half * 32;
          if (l < 4) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                float x = v[hh][c] + b[hh * 16 + c];
                if (l == 2) x += tab[hh * 16 + c];                     // o1 = relu(tables + W1 [h2 | ang] + b1)
                x = fmaxf(x, 0.f);
                if (l == 1) x *= keepf;                                // h2 * structure pair mask (pair.py:85-87)
                v[hh][c] = x;
              }
            pt_store16(t_act, PT_TM_ACTLO - PT_TM_ACT, v[0]);
            pt_store16(t_act + 16, PT_TM_ACTLO - PT_TM_ACT, v[1]);
            if (l == 1) pt_store16(tlane + PT_TM_ACT + 64 + half * 16, PT_TM_ACTLO - PT_TM_ACT, ang);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(act_ready);
          } else {
            // last layer: (W3 o2 + b3) * has_CA pair mask, 128 contiguous bytes per thread
            const int j = j0 + p;
            if (j < L) {
              const float s = sOk[p] ? 1.f : 0.f;                                        // pair.py:99
              float4* dst = reinterpret_cast<float4*>(a.out + ((size_t)row * L + j) * 64 + half * 32);
#pragma unroll
              for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int c = 0; c < 16; c += 4)
                  __stcs(dst + hh * 4 + (c >> 2), make_float4((v[hh][c] + b[hh * 16 + c]) * s, (v[hh][c + 1] + b[hh * 16 + c + 1]) * s,
                                                               (v[hh][c + 2] + b[hh * 16 + c + 2]) * s, (v[hh][c + 3] + b[hh * 16 + c + 3]) * s));
            }
            tc_fence_before();
          }
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
}

}  // namespace
}  // namespace abopt

using namespace abopt;

// ------------------------------------------------------------------------------------------ C ABI
struct abopt_pair_embed {
  int device = 0, A = 0;
  bool finalized = false;
  std::map<std::string, size_t> spec;
  std::map<std::string, std::vector<float>> sd;
  void* wbase = nullptr;
  PairEmbedW w;
  int sm_count = 148;
};

#define PE_CUDA_TRY(expr)                                                                       \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) return api_fail(ABOPT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)

extern "C" int abopt_pair_embed_create(int max_num_atoms, int device, abopt_pair_embed** out) {
  if (!out) return api_fail(ABOPT_ERR_ARG, "null argument");
  if (max_num_atoms < 4 || max_num_atoms > PE_MAXA) return api_fail(ABOPT_ERR_ARG, "max_num_atoms must be in [4, 15] (backbone N, CA, C, O at least)");
  int ndev = 0;
  PE_CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return api_fail(ABOPT_ERR_ARG, "no such CUDA device");
  cudaDeviceProp prop;
  PE_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return api_fail(ABOPT_ERR_CUDA, std::string("libabopt_b200 needs a B200-class GPU (sm_100); found ") + prop.name);
  abopt_pair_embed* pe = new abopt_pair_embed();
  pe->device = device; pe->A = max_num_atoms; pe->sm_count = prop.multiProcessorCount;
  const size_t A2 = (size_t)max_num_atoms * max_num_atoms;
  pe->spec = {{"aa_pair_embed.weight", (size_t)PE_AA * PE_AA * 64}, {"relpos_embed.weight", (size_t)(2 * PE_RELPOS + 1) * 64},
              {"aapair_to_distcoef.weight", (size_t)PE_AA * PE_AA * A2}, {"dihedral_embed.freq_bands", 6},
              {"distance_embed.0.weight", 64 * A2}, {"distance_embed.0.bias", 64},
              {"distance_embed.2.weight", 64 * 64}, {"distance_embed.2.bias", 64},
              {"out_mlp.0.weight", (size_t)64 * (192 + PE_ANG)}, {"out_mlp.0.bias", 64},
              {"out_mlp.2.weight", 64 * 64}, {"out_mlp.2.bias", 64}, {"out_mlp.4.weight", 64 * 64}, {"out_mlp.4.bias", 64}};
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(device);
  cudaError_t e = cudaFuncSetAttribute(pair_embed_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PT_SMEM);
  cudaSetDevice(cur);
  if (e != cudaSuccess) { delete pe; return api_fail(ABOPT_ERR_CUDA, std::string("pair_embed_tc_kernel shared memory: ") + cudaGetErrorString(e)); }
  *out = pe;
  return ABOPT_OK;
}

extern "C" void abopt_pair_embed_destroy(abopt_pair_embed* pe) {
  if (!pe) return;
  if (pe->wbase) {
    int cur = 0;
    cudaGetDevice(&cur); cudaSetDevice(pe->device);
    cudaFree(pe->wbase);
    cudaSetDevice(cur);
  }
  delete pe;
}

extern "C" int abopt_pair_embed_set_tensor(abopt_pair_embed* pe, const char* key, const float* data, size_t numel, int on_device) {
  if (!pe || !key) return api_fail(ABOPT_ERR_ARG, "null argument");
  auto it = pe->spec.find(key);
  if (it == pe->spec.end()) return api_fail(ABOPT_ERR_KEY, std::string("unexpected state-dict key: ") + key);
  if (it->second != numel) return api_fail(ABOPT_ERR_KEY, std::string("size mismatch for ") + key + ": expected " + std::to_string(it->second) +
                                                              " elements, got " + std::to_string(numel));
  if (!data) return api_fail(ABOPT_ERR_ARG, "null data");
  std::vector<float>& t = pe->sd[key];
  t.resize(numel);
  if (on_device) {
    int cur = 0;
    cudaGetDevice(&cur); cudaSetDevice(pe->device);
    cudaError_t e = cudaMemcpy(t.data(), data, numel * sizeof(float), cudaMemcpyDeviceToHost);
    cudaSetDevice(cur);
    if (e != cudaSuccess) return api_fail(ABOPT_ERR_CUDA, std::string("cudaMemcpy: ") + cudaGetErrorString(e));
  } else {
    memcpy(t.data(), data, numel * sizeof(float));
  }
  pe->finalized = false;
  return ABOPT_OK;
}

extern "C" int abopt_pair_embed_finalize(abopt_pair_embed* pe) {
  if (!pe) return api_fail(ABOPT_ERR_ARG, "null argument");
  for (auto& kv : pe->spec)
    if (!pe->sd.count(kv.first)) return api_fail(ABOPT_ERR_STATE, "missing state-dict key: " + kv.first);
  const int A2 = pe->A * pe->A, NP = PE_AA * PE_AA, NR = 2 * PE_RELPOS + 1, K1 = 192 + PE_ANG;
  const std::vector<float>& W1 = pe->sd["out_mlp.0.weight"];
  std::vector<float> img;
  auto reserve = [&](size_t n) { size_t off = (img.size() + 63) & ~size_t(63); img.resize(off + n, 0.f); return off; };
  const size_t o_taa = reserve((size_t)NP * 64), o_trel = reserve((size_t)NR * 64), o_bias = reserve(5 * 64);
  std::vector<float> coef((size_t)NP * A2);
  {  // F.softplus (beta 1, threshold 20), pair.py:81
    const std::vector<float>& c = pe->sd["aapair_to_distcoef.weight"];
    for (size_t k = 0; k < c.size(); ++k) coef[k] = c[k] > 20.f ? c[k] : log1pf(expf(c[k]));
  }
  auto premul = [&](const std::vector<float>& E, int rows, int col0, size_t off) {      // T[r][o] = sum_c E[r][c] W1[o][col0 + c]
    for (int r = 0; r < rows; ++r)
      for (int o = 0; o < 64; ++o) {
        double s = 0.0;
        for (int c = 0; c < 64; ++c) s += (double)E[(size_t)r * 64 + c] * (double)W1[(size_t)o * K1 + col0 + c];
        img[off + (size_t)r * 64 + o] = (float)s;
      }
  };
  premul(pe->sd["aa_pair_embed.weight"], NP, 0, o_taa);
  premul(pe->sd["relpos_embed.weight"], NR, 64, o_trel);
  // ---- pre-swizzled K-major weight boxes, hi | lo planes
  const int kb1 = (pe->A + 1) / 2;                        // layer-1 K order: [key atom][query atom padded to 16], 2 key atoms per k-block
  const size_t o_box = reserve((size_t)(kb1 + 9) * 4096);
  const size_t o_coefT = reserve((size_t)NP * pe->A * 16);
  for (int r = 0; r < NP; ++r)
    for (int ib = 0; ib < pe->A; ++ib)
      for (int ia = 0; ia < pe->A; ++ia) img[o_coefT + ((size_t)r * pe->A + ib) * 16 + ia] = coef[(size_t)r * A2 + ia * pe->A + ib];
  auto pack_box = [&](size_t box, auto val) {             // val(k, n) for k in [0, 32), n in [0, 64)
    float* hi = &img[o_box + box * 4096];
    float* lo = hi + 2048;
    for (int n = 0; n < 64; ++n)
      for (int k = 0; k < 32; ++k) {
        const float v = val(k, n);
        uint32_t bits;
        memcpy(&bits, &v, 4);
        bits &= 0xFFFFE000u;
        float h;
        memcpy(&h, &bits, 4);
        const size_t at = (size_t)n * 32 + (size_t)(((k >> 2) ^ (n & 7)) << 2) + (k & 3);
        hi[at] = v;                                        // raw fp32: the tensor core ignores the low 13 mantissa bits
        lo[at] = v - h;
      }
  };
  {
    const std::vector<float>& Wd1 = pe->sd["distance_embed.0.weight"];      // [64][A2]
    for (int kb = 0; kb < kb1; ++kb)
      pack_box(kb, [&](int k, int n) {
        const int ib = 2 * kb + (k >> 4), ia = k & 15;
        return (ib < pe->A && ia < pe->A) ? Wd1[(size_t)n * A2 + ia * pe->A + ib] : 0.f;
      });
    const std::vector<float>& Wd2 = pe->sd["distance_embed.2.weight"];
    const std::vector<float>& W2 = pe->sd["out_mlp.2.weight"];
    const std::vector<float>& W3 = pe->sd["out_mlp.4.weight"];
    for (int kb = 0; kb < 2; ++kb) {
      pack_box(kb1 + 0 + kb, [&](int k, int n) { return Wd2[(size_t)n * 64 + kb * 32 + k]; });
      pack_box(kb1 + 2 + kb, [&](int k, int n) { return W1[(size_t)n * K1 + 128 + kb * 32 + k]; });
      pack_box(kb1 + 5 + kb, [&](int k, int n) { return W2[(size_t)n * 64 + kb * 32 + k]; });
      pack_box(kb1 + 7 + kb, [&](int k, int n) { return W3[(size_t)n * 64 + kb * 32 + k]; });
    }
    pack_box(kb1 + 4, [&](int k, int n) {
      const int which = k >> 4, idx = k & 15;
      return idx < 13 ? W1[(size_t)n * K1 + 192 + which * 13 + idx] : 0.f;
    });
  }
  const char* bkeys[5] = {"distance_embed.0.bias", "distance_embed.2.bias", "out_mlp.0.bias", "out_mlp.2.bias", "out_mlp.4.bias"};
  for (int b = 0; b < 5; ++b) memcpy(&img[o_bias + b * 64], pe->sd[bkeys[b]].data(), 64 * sizeof(float));

  int cur = 0;
  cudaGetDevice(&cur); cudaSetDevice(pe->device);
  if (pe->wbase) { cudaFree(pe->wbase); pe->wbase = nullptr; }
  cudaError_t e = cudaMalloc(&pe->wbase, img.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(pe->wbase, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice);
  cudaSetDevice(cur);
  if (e != cudaSuccess) return api_fail(ABOPT_ERR_CUDA, std::string("pair-embed weights: ") + cudaGetErrorString(e));
  const float* base = static_cast<const float*>(pe->wbase);
  pe->w.A = pe->A; pe->w.A2 = A2;
  pe->w.Taa = base + o_taa; pe->w.Trel = base + o_trel; pe->w.bias = base + o_bias;
  pe->w.boxes = base + o_box; pe->w.kb1 = kb1; pe->w.coefT = base + o_coefT;
  memcpy(pe->w.freq, pe->sd["dihedral_embed.freq_bands"].data(), 6 * sizeof(float));
  pe->finalized = true;
  return ABOPT_OK;
}

extern "C" int abopt_pair_embed_forward(abopt_pair_embed* pe, int N, int L, int num_atoms_in, const int64_t* aa, const int64_t* res_nb,
                                        const int64_t* chain_nb, const float* pos_atoms, const uint8_t* mask_atoms,
                                        const uint8_t* structure_mask, const uint8_t* sequence_mask, float* pair_feat, void* stream) {
  if (!pe) return api_fail(ABOPT_ERR_ARG, "null handle");
  if (!pe->finalized) return api_fail(ABOPT_ERR_STATE, "pair embedding not finalised");
  if (N < 0 || L < 0) return api_fail(ABOPT_ERR_ARG, "negative size");
  if (num_atoms_in < pe->A) return api_fail(ABOPT_ERR_ARG, "pos_atoms / mask_atoms have fewer atoms per residue than max_num_atoms");
  if (N == 0 || L == 0) return ABOPT_OK;
  if (!aa || !res_nb || !chain_nb || !pos_atoms || !mask_atoms || !pair_feat) return api_fail(ABOPT_ERR_ARG, "null tensor");
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != pe->device) cudaSetDevice(pe->device);
  PairEmbedArgs a{N, L, num_atoms_in, (const long long*)aa, (const long long*)res_nb, (const long long*)chain_nb, pos_atoms, mask_atoms,
                  structure_mask, sequence_mask, pair_feat};
  const long long rows = (long long)N * L;
  const int grid = (int)std::min<long long>(rows, pe->sm_count);
  cudaStream_t st = (cudaStream_t)stream;
  {
    ProfScope ps(KK_OTHER, st);
    pair_embed_tc_kernel<<<grid, PT_THREADS, PT_SMEM, st>>>(pe->w, a);
  }
  cudaError_t e = cudaGetLastError();
  if (cur != pe->device) cudaSetDevice(cur);
  if (e != cudaSuccess) return api_fail(ABOPT_ERR_CUDA, std::string("pair_embed_tc_kernel: ") + cudaGetErrorString(e));
  return ABOPT_OK;
}
