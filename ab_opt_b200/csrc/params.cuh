// Packed device-side weight views handed to the kernels (built by abopt_model_finalize).
#pragma once
#include "common.cuh"

namespace abopt {

// One GABlock (modules/encoders/ga.py:41-79).  "t" suffix = stored K-major (transposed nn.Linear weight).
struct BlockW {
  const float* Wcat;       // [2016][128]  rows: proj_query | proj_key | proj_value | proj_query_point | proj_key_point | proj_value_point
  const float* Wcat_lo;    // [2016][128]  tf32 "lo" plane of Wcat (3xTF32 tensor-core GEMM)
  const float* Wout;       // [128][1824]  out_transform.weight as stored by nn.Linear (K-major B operand)
  const float* Wout_lo;    // [128][1824]
  const float* Wb;         // [64][12]     proj_pair_bias.weight transposed (c-major)
  const float* coef;       // [12]         -softplus(spatial_coef) * sqrt(2/(9*8)) / 2
  const float* Wout_t;     // [1824][128]  out_transform.weight^T
  const float* bout;       // [128]
  const float* ln1_g; const float* ln1_b;
  const float* W1_t; const float* b1;      // mlp_transition.0  [128][128]^T
  const float* W2_t; const float* b2;      // mlp_transition.2
  const float* W3_t; const float* b3;      // mlp_transition.4
  const float* Wmlp;       // [3 * 128][128]  mlp_transition.{0,2,4}.weight stacked as stored by nn.Linear (K-major B operands)
  const float* Wmlp_lo;    // [3 * 128][128]  tf32 lo plane
  const float* ln2_g; const float* ln2_b;
  const float* Wb_raw;     // [12][64]     proj_pair_bias.weight as stored (backward pass)
  const float* sc_raw;     // [12]         spatial_coef as stored (backward pass: sigmoid)
};

// pair-bias weight + spatial coefficients as BY-VALUE kernel parameters: they land in the constant
// bank, so the 768 FFMAs per residue pair take them as immediate c[0x0][..] operands (no loads).
struct PairBiasParams {
  float Wb[C][H];          // 3072 B
  float coef[H];
};

// The same pair-bias weight packed as head PAIRS for the FFMA2 (fma.rn.f32x2) inner loop of pair_bias_kernel:
// w[half][c][k] = (W_b[6 half + 2k][c], W_b[6 half + 2k + 1][c]); by-value kernel parameter -> constant bank ->
// uniform registers (the head half is warp-uniform in the kernel).
struct PairBiasPacked {
  float2 w[2][C][3];       // 3072 B
};

// One 3-layer head (eps_crd_net / eps_rot_net / eps_seq_net, dpm_full.py:45-62).
struct HeadW {
  const float* W0_t;       // [128][128]   first 128 input columns of layer 0, transposed
  const float* W0_ext;     // [3][128]     columns 128..130 (beta, sin beta, cos beta), transposed
  const float* b0;
  const float* W2_t; const float* b2;
  const float* W4_t;       // [128][128]   zero-padded to 128 output columns
  const float* b4;         // [128]        zero-padded
};

struct EpsW {
  const float* emb;        // [25][128]   current_sequence_embedding.weight
  const float* Wm0_t;      // [256][128]  res_feat_mixer.0^T
  const float* bm0;
  const float* Wm2_t;      // [128][128]
  const float* bm2;
  HeadW crd, rot, seq, prm;    // prm = prmsd_predictor (linear_1..3), valid iff has_prmsd
  const float* prm_ln_g;   // [131]
  const float* prm_ln_b;   // [131]
  int has_prmsd;
  int prmsd_bins;
};

// Diffusion constants (modules/diffusion/transition.py:10-34, modules/common/so3.py:70-109)
struct DiffW {
  const float* betas; const float* alpha_bars; const float* alphas; const float* sigmas;     // trans_pos.var_sched
  const float* alpha_bars_rot;   // trans_rot.var_sched.alpha_bars (add_noise)
  const float* alpha_bars_seq;   // trans_seq.var_sched.alpha_bars (posterior / add_noise)
  const float* sqrt_recip_ab; const float* sqrt_recipm1_ab;      // each [T+1]
  const float* ang_X[2];        // [T+1][8192]  (0 = fwd, 1 = inv)
  const float* ang_Y[2];
  const float* ang_cdf[2];      // [T+1][8192]  normalised inclusive cumsum of Y[:, :-1] (fast mode)
  const float* ang_std[2];      // [T+1]
  const uint8_t* ang_flag[2];   // [T+1]
  float pos_mean[3]; float pos_scale;
  int num_steps;
  int obj_pred_x0;
  float prmsd_min, prmsd_max;
};

}  // namespace abopt
