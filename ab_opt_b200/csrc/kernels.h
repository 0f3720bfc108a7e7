// Host-side launcher declarations shared by api.cu and the kernel translation units.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "params.cuh"

namespace abopt {

struct NoisePtrs {
  const float* u; const float* unif_ang; const float* gauss_ang; const float* z_pos; const float* expo_seq;
  const int* bin_idx;
};
struct StepArgs {
  int M, L, t;
  int sample_structure, sample_sequence, pred_x0, masked_ppl;
  const float* v_t; const float* p_t_ang; const long long* s_t;
  const float* v_net; const float* p_pred; const float* c_den;
  const uint8_t* mask_gen;
  float* v_out; float* p_out_ang; long long* s_out;
  float* maxprob_rows;
  NoisePtrs nz;
  uint64_t seed;
  uint32_t row0;                // first row of this batch in the global (unsharded) batch: Philox counters are global rows
};
struct InitArgs {
  int M, L, T0;
  int sample_structure, sample_sequence, optimize, has_prmsd;
  const float* v; const float* p_ang; const long long* s; const uint8_t* mask_gen;
  const float* g4; const float* gp; const long long* s_rand;
  NoisePtrs add;
  float* v_out; float* p_out_ang; long long* s_out;
  float* prmsd_out; float* ppl_out;
  uint64_t seed;
  uint32_t row0;                // see StepArgs
  // FullDPM.forward (training): per-complex steps, the sequence draw on every row, the position noise kept
  const long long* tvec;        // (N,) steps; nullptr -> T0 for every complex
  int seq_all_rows;             // 1: _sample(c_t) also where nothing is generated (transition.py:198-199)
  float* z_out;                 // (M,3) the position noise e_rand (transition.py:74), may be nullptr
  int grad_clamp;               // 1: log_rotation as evaluated with autograd enabled (cosine clamped at -0.999, so3.py:12-17)
};
// FullDPM.forward losses (dpm_full.py:156-234; AbDesign :138-190)
struct LossArgs {
  int N, L, abdock, pred_x0, has_prmsd, bins;
  float dmin, dmax;
  const float* v_0; const float* p_0_ang; const long long* s_0;
  const float* p_noisy_ang; const long long* s_noisy; const float* z;      // z may be nullptr (denoise_structure = False)
  const float* R_pred; const float* p_pred; const float* c_den; const float* prmsd_logits;
  const uint8_t* mask_gen; const uint8_t* mask_res; const long long* tvec;
  float* rows;                  // scratch [6][M]: rot | pos | seq | squared deviation (A^2) | dist sum | dist count
  float* out;                   // [5]: rot, pos, seq, prmsd, dist
};
void launch_loss(const LossArgs& a, const DiffW& dw, cudaStream_t st);
void launch_gather_beta(int N, const long long* tvec, const float* betas, float* out, cudaStream_t st);

cudaError_t linear_kernels_init();
cudaError_t attn_kernels_init();
cudaError_t tc_init();
void launch_split(const float* in, float* hi, float* lo, size_t n, cudaStream_t st);
void launch_lo(const float* in, float* lo, size_t n, cudaStream_t st);
void launch_transpose_split(const float* W, int R, int Cc, float* Th, float* Tl, cudaStream_t st);
bool launch_gemm3x_plain(int M, int N, int K, const float* Ah, const float* Al, int lda, const float* Bh, const float* Bl, int ldb,
                         float* D, int ldd, const float* bias, cudaStream_t st);

cudaError_t tail_tc_init();
void tail_debug_clocks(long long* out6);
bool launch_outT_tail(int M, const float* feat, const float* x, const uint8_t* mask, const BlockW& w, float* x_out,
                      float* x_lo_out, cudaStream_t st, const int* count = nullptr);

// operands of the tensor-core attention kernels, produced by the projection GEMM epilogue (k_tc.cu: EpiProjPack)
struct AttnOperands {
  float* QA; float* QA_lo;      // [N][H][L][64]
  float* KB; float* KB_lo;      // [N][H][L][64]
  float* rq; float* rk;         // [N][H][L]
  float* VT;                    // [N][H][64][Lp]  values, key index contiguous (rows 56..63 and columns >= L stay zero)
};
bool launch_proj_pack(int M, int L, int Lp, const float* xh, const float* xl, const float* Wh, const float* Wl, const float* R, const float* t,
                      const float* coef, const AttnOperands& op, cudaStream_t st, const int* rows = nullptr, const int* count = nullptr);
cudaError_t aggr_tc_init();
bool launch_aggr_tc(int nb, int b0, int N, int L, int Lp, const float* alpha, const float* VT,
                    const float* R, const float* t, float* feat, cudaStream_t st, const int2* windows = nullptr,
                    const int* wcount = nullptr, const int* cidx = nullptr);
bool make_tmap(CUtensorMap* m, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
bool make_tmap_plain(CUtensorMap* m, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols);
cudaError_t attn_tc_init();
bool attn_needs_qk_lo(int L);      // true when the logits kernel for this length reads QA_lo / KB_lo from global memory
void attn_debug_clocks(long long* out16);
// final logits + softmax on the tensor cores: alpha[chunk][h][i][Lp] for complexes [b0, b0 + nb)
bool launch_attn_logits_tc(int nb, int b0, int N, int L, int Lp, const AttnOperands& op, const float* bias_layer, const uint8_t* mask,
                           float* alpha, cudaStream_t st, const int2* windows = nullptr, const int* wcount = nullptr,
                           const uint8_t* exclude = nullptr, float2* stats = nullptr);
bool make_tmap_3d(CUtensorMap* m, const float* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box0, uint32_t box1);

void launch_mixer(int M, const float* res_feat, const long long* s_t, const float* v_t, const EpsW& w,
                  float* x_out, float* Rbuf, const float* p_ang, float* p_norm, const float* mean, float scale,
                  float* x_lo_out, cudaStream_t st, const int* rows = nullptr, const int* count = nullptr, float* x_c = nullptr,
                  float* x_c_lo = nullptr);
void launch_heads(int M, int L, const float* x, const float* beta, int beta_stride, const float* Rbuf, const float* v_t,
                  const uint8_t* mask_gen, const EpsW& w, float* v_next, float* R_next, float* eps_pos, float* c_den,
                  float* prmsd_rows, float* prmsd_logits, cudaStream_t st, const int* rows = nullptr, const int* count = nullptr,
                  bool x_compact = false);
// focus mode (k_linear.cu): restriction of the last GABlock to the generated rows inside the sampling loop
struct Focus {
  int* cidx;          // [M]  compact row of residue row r, -1 = not needed
  int* rows;          // [M]  residue row of compact row k
  int2* windows;      // [N * ceil(L / 128) + N]  (complex, first query row) of the 128-row query windows
  int* count;         // [2]  count[0] = needed rows, count[1] = windows (device)
  int* scratch;       // [2 N]
  float* x_c;         // [M][128] gathered block input of the compact rows
  uint8_t* mask_c;    // [M]
};
void launch_focus_build(int N, int L, const uint8_t* mask_gen, int* cidx, int* rows, int2* windows, int* count, int* scratch,
                        cudaStream_t st);
void launch_focus_gather(const int* rows, const int* count, const float* x, const uint8_t* mask, float* x_c, uint8_t* mask_c,
                         cudaStream_t st);

void launch_alpha_tap(int nb, int b0, int L, int Lp, const float* alpha, float* out, cudaStream_t st);
// k_pair.cu: pair_bias_kernel (hoisted z . W_b) and pair_stream_kernel (streams z once per GABlock)
cudaError_t pair_stream_init();
bool make_tmap_2d(CUtensorMap* m, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols);
bool launch_pair_bias(int nb, int b0, int N, int L, int Lp, const float* z, const PairBiasPacked& pb, float* bias, cudaStream_t st);
// row list of one pair_stream launch (pair_rows_build_kernel): live rows from the front, masked-but-needed rows from the back
struct PairRows {
  int4* list;         // [nb * L]  (complex within the launch, residue, output row, -)
  int* count;         // [2]  live rows | masked rows that still need their zeros (device)
};
// cidx (focus mode): compact output row of residue row r, -1 = row not needed at all
// compact: the output row of a listed row is cidx[row] (focus mode); otherwise the row itself
void launch_pair_rows_build(int nb, int b0, int L, const uint8_t* mask, const int* cidx, const PairRows& pr, cudaStream_t st,
                            bool compact = true);
bool launch_pair_stream(int nb, int b0, int L, int Lp, const float* z, float* alpha, float* feat, const PairRows& pr, cudaStream_t st,
                        int feat_ld = NFEAT, bool partial = false);      // partial: the list holds the generated rows only (profile kind)
// context cache of the first GABlock inside the sampling loop (k_pair.cu: ctx_delta_kernel)
void launch_ctx_delta(int N, int L, int Lp, const float* z, const uint8_t* mask, float* alpha, const float* cache,
                      const float2* stats_ctx, const float2* stats, const int* cidx, const int* rows, const int* first,
                      const int* count, float* feat, cudaStream_t st);
bool make_tmap_3d_plain(CUtensorMap* m, const float* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box0, uint32_t box1, uint32_t box2);
bool make_tmap_3d_sw128(CUtensorMap* m, const float* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box0, uint32_t box1, uint32_t box2);

void launch_angle_argmax(int M, int L, const long long* tvec, int t_uniform, const float* Y, const float* expo,
                         const uint8_t* mask_gen, int* bin_idx, cudaStream_t st);
void launch_step(const StepArgs& a, const DiffW& dw, cudaStream_t st);
void launch_complex_reduce(int N, int L, int bins, float dmin, float dmax, int masked_ppl, const float* prmsd_logits,
                           const float* maxprob_rows, const uint8_t* mask_gen, float* prmsd_out, float* ppl_out, cudaStream_t st);
void launch_init(const InitArgs& a, const DiffW& dw, cudaStream_t st);
void launch_design_prep(int M, int A_in, const float* pos, const uint8_t* mask_atoms, const uint8_t* gen, uint8_t* ctx, float* v0,
                        float* p0, cudaStream_t st);
void launch_rot_denoise(int M, int L, const float* v_t, const float* v_net, const uint8_t* mask_gen, const long long* tvec,
                        const NoisePtrs& nz, const DiffW& dw, float* v_out, cudaStream_t st);
void launch_pos(int M, int L, int mode, const float* p_t, const float* other, const uint8_t* mask_gen, const long long* tvec,
                const float* z_pos, const DiffW& dw, float* out, cudaStream_t st);
void launch_seq_denoise(int M, int L, const long long* s_t, const float* c0, const uint8_t* mask_gen, const long long* tvec,
                        const float* expo_seq, const DiffW& dw, float* post, long long* s_out, cudaStream_t st);

// ---- k_backward.cu: the backward pass of the training step (fp32 CUDA-core GEMMs + row / pair kernels)
cudaError_t backward_kernels_init();
void bwd_gemm_nt(int M, int N, int K, const float* x, int ldx, bool relu_x, const float* W, int ldw, const float* bias, float* y, int ldy,
                 cudaStream_t st);
void bwd_gemm_nn(int M, int N, int K, const float* g, int ldg, const float* W, int ldw, float* dx, int lddx, bool accumulate, cudaStream_t st);
void bwd_wgrad(int M, int N, int K, const float* g, int ldg, const float* x, int ldx, bool relu_x, float* dW, float* scratch,
               size_t scratch_floats, cudaStream_t st);
void bwd_colsum(int M, int K, const float* A, int lda, const float* B, int ldb, float* out, float* scratch, cudaStream_t st, float scale);
void bwd_add_ln_fwd(int M, const float* a, const float* b, const uint8_t* mask, const float* gamma, const float* beta, float* s_out,
                    float* h_out, cudaStream_t st);
void bwd_ln(int M, const float* g, const float* g2, const float* s, const float* gamma, float* ds, float* gxh, float* gsum, cudaStream_t st);
void bwd_relu(size_t n, float* g, const float* a, cudaStream_t st);
void bwd_add_mask(int M, int K, const float* a, const float* b, const uint8_t* mask, float* out, cudaStream_t st);
void bwd_points_global(int M, const float* Pm, const float* R, const float* t, float* PG, cudaStream_t st);
void bwd_aggregate(int M, const float* gfeat, const float* feat, const float* R, float* g_agg, cudaStream_t st);
struct PairBwdArgs {
  int N, L, Lp;
  const float* z; const float* alpha; const uint8_t* mask;
  const float* gfeat;            // (M, 1824): cols 0..767 = g_p2n, 768..1151 = g_node
  const float* g_agg;            // (M, 288)
  const float* Pm;               // (M, 2016) plain projections (q | k | v | local points)
  const float* PG;               // (M, 864) global points q | k | v
  const float* R;                // (M, 9)
  const float* Wb;               // [12][64] proj_pair_bias.weight
  const float* coef;             // [12]
  float* g_log;                  // [N][H][L][Lp]
  float* G;                      // (M, 2016) gradient of the plain projections
  float* dz; int dz_accumulate;  // (N, L, L, 64)
  float* part;                   // (M, 780): per-row partial d W_b (768) | d coef (12)
};
void launch_pair_bwd(const PairBwdArgs& a, cudaStream_t st);
void bwd_mixer_cat(int M, const float* res_feat, const long long* s_t, const float* emb, float* cat0, cudaStream_t st);
void bwd_heads_cat(int M, int L, const float* x, const float* beta, float* hcat, cudaStream_t st);
void bwd_embed_grad(int M, const long long* s_t, const float* g_cat, float* dE, cudaStream_t st);
void bwd_ln131_fwd(int M, const float* h, const float* gamma, const float* beta, float* out, cudaStream_t st);
void bwd_ln131_bwd(int M, const float* g, const float* h, const float* gamma, float* dh_acc, float* gxh, cudaStream_t st);
struct LossBwdArgs {
  int N, L, abdock, pred_x0, has_prmsd, bins;
  float dmin, dmax;
  float lw[5];                   // loss weights: rot, pos, seq, prmsd, dist
  const float* v_0; const float* p_0_ang; const long long* s_0;
  const float* p_noisy_ang; const long long* s_noisy; const float* z;
  const float* R; const float* R_pred; const float* eps_pos; const float* c_den;
  const float* o_rot; int ld_orot;          // raw output of eps_rot_net (M, 3)
  const float* prmsd_logits;
  const uint8_t* mask_gen; const uint8_t* mask_res; const long long* tvec;
  const float* rows;             // [6][M] per-residue terms written by the forward loss kernel
  float* stats;                  // [4]: n_gen + 1e-8 | dist count | sum mask_generate[:, 0] + 1e-10
  float* glog;                   // (N, bins) d prmsd loss / d logits / L
  float* GO;                     // (M, 26): d loss / d (eps_crd_net | eps_rot_net | eps_seq_net) outputs
  float* GP;                     // (M, bins) d loss / d prmsd head rows (may be null)
};
void launch_loss_bwd(const LossBwdArgs& a, const DiffW& dw, cudaStream_t st);

}  // namespace abopt
