// Per-residue diffusion transitions, one thread per residue, everything in registers:
//   RotationTransition.denoise / add_noise        modules/diffusion/transition.py:120-160, common/so3.py:111-146
//   PositionTransition.denoise / add_noise / pred_noise_from_start        transition.py:42-101
//   AminoacidCategoricalTransition.denoise / add_noise                    transition.py:170-245
// plus the initialisation of FullDPM.sample / optimize and the per-complex pRMSD / perplexity
// reductions (dpm_full.py:255-267, 284, 293, 321-337, 380-399).
//
// Randomness is either replayed from caller-supplied draws ("parity" mode: the same ATen draws the
// reference consumes, so sampled indices are bit-identical) or generated in-kernel with Philox4x32-10
// keyed by (seed; residue, step, stream) ("fast" mode).
#include "kernels.h"

namespace abopt {

// The categorical maths must round exactly like the reference's unfused ATen ops so that the sampled
// amino-acid index is bit-identical: no FMA contraction here.
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float div_(float a, float b) { return __fdiv_rn(a, b); }

enum RngStream : uint32_t { RS_U = 1, RS_ANGLE = 2, RS_ZPOS = 3, RS_SEQ = 4, RS_INIT_G4 = 10, RS_INIT_GP = 11, RS_INIT_S = 12 };


// ---------------------------------------------------------------- angle sampling (so3.py:111-138)
// parity: bin = argmax_k Y[t][k] / q[r][k], first index on ties (== torch.multinomial(prob[:, :-1], 1))
__global__ void angle_argmax_kernel(int M, int L, const long long* __restrict__ tvec, int t_uniform,
                                    const float* __restrict__ Y, const float* __restrict__ expo,
                                    const uint8_t* __restrict__ mask_gen, int* __restrict__ bin_idx) {
  const int r = blockIdx.x;
  if (r >= M) return;
  if (mask_gen != nullptr && mask_gen[r] == 0) { if (threadIdx.x == 0) bin_idx[r] = 0; return; }   // result is discarded
  const int t = tvec ? (int)tvec[r / L] : t_uniform;
  const float* y = Y + (size_t)t * NBINS;
  const float* q = expo + (size_t)r * (NBINS - 1);
  float best = -INFINITY; int bi = 0x7fffffff;
  for (int k = threadIdx.x; k < NBINS - 1; k += blockDim.x) {
    const float v = div_(y[k], q[k]);
    if (v > best || (v == best && k < bi)) { best = v; bi = k; }
  }
  __shared__ float sb[32]; __shared__ int si[32];
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sb[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (sb[w] > best || (sb[w] == best && si[w] < bi)) { best = sb[w]; bi = si[w]; }
    bin_idx[r] = bi;
  }
}

// fast: inverse CDF over the same histogram (first bin whose cumulative mass exceeds u)
__device__ __forceinline__ int cdf_search(const float* __restrict__ cdf, float u) {
  int lo = 0, hi = NBINS - 2;                 // bins 0 .. 8190
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cdf[mid] > u) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// theta ~ angular distribution `tab` at schedule index t
__device__ __forceinline__ float sample_angle(const DiffW& dw, int tab, int t, int bin_parity, float u_cdf,
                                              float unif, float gauss, bool parity) {
  const float sd = dw.ang_std[tab][t];
  float theta;
  if (dw.ang_flag[tab][t]) {                                       // Gaussian approximation, so3.py:129-132
    theta = fmodf(fabsf(add_(mul_(sd, 2.f), mul_(gauss, sd))), 3.14159274101257324f);
  } else {                                                          // histogram, so3.py:122-126
    const int bin = parity ? bin_parity : cdf_search(dw.ang_cdf[tab] + (size_t)t * NBINS, u_cdf);
    const float* X = dw.ang_X[tab] + (size_t)t * NBINS;
    const float start = X[bin], width = X[bin + 1] - X[bin];
    theta = add_(start, mul_(unif, width));
  }
  return theta;
}

// e = normalize(u) * theta  (so3.py:141-146);  returns log( exp(e) * exp(base) )
__device__ __forceinline__ void compose_noise_rotation(float ux, float uy, float uz, float theta, bool add_noise,
                                                       float bx, float by, float bz, float& ox, float& oy, float& oz, float lo = -1.f) {
  const float n = fmaxf(sqrtf(ux * ux + uy * uy + uz * uz), 1e-12f);     // F.normalize eps
  float ex = ux / n * theta, ey = uy / n * theta, ez = uz / n * theta;
  if (!add_noise) { ex = 0.f; ey = 0.f; ez = 0.f; }                        // transition.py:149-153
  const Mat3 E = so3_exp(ex, ey, ez);
  const Mat3 B = so3_exp(bx, by, bz);
  const Mat3 Rn = matmul3(E, B);
  so3_log(Rn, ox, oy, oz, lo);
}

// ---------------------------------------------------------------- categorical posterior + sample
// post_k (transition.py:202-227,239-240), returns argmax_k (post_k + 1e-8) / q_k (transition.py:176-180)
__device__ __forceinline__ int seq_posterior_sample(long long s_t, const float* c0, float ab, bool gen,
                                                    const float* q, float* post_out, float& maxprob) {
  const float base = div_(add_(1.f, -ab), (float)NAA);      // (1 - alpha_bar) / K
  float th[NAA], sum = 0.f;
  const bool s_ok = (s_t >= 0 && s_t < NAA);
#pragma unroll
  for (int k = 0; k < NAA; ++k) {
    const float ct = (s_ok && k == (int)s_t) ? 1.f : 0.f;  // clampped_one_hot, layers.py:10-14
    th[k] = mul_(add_(mul_(ab, ct), base), add_(mul_(ab, c0[k]), base));
    sum = add_(sum, th[k]);
  }
  const float den = add_(sum, 1e-8f);
  float best = -INFINITY; int bi = 0;
  float pm = -INFINITY;
#pragma unroll
  for (int k = 0; k < NAA; ++k) {
    const float ct = (s_ok && k == (int)s_t) ? 1.f : 0.f;
    const float pk = gen ? div_(th[k], den) : ct;
    th[k] = pk;
    if (post_out) post_out[k] = pk;
    pm = fmaxf(pm, pk);
    const float v = div_(add_(pk, 1e-8f), q[k]);
    if (v > best) { best = v; bi = k; }
  }
  // max softmax probability of `post` (calc_perplexity, dpm_full.py:393)
  float se = 0.f;
#pragma unroll
  for (int k = 0; k < NAA; ++k) se += expf(th[k] - pm);
  maxprob = 1.f / se;
  return bi;
}

__device__ __forceinline__ void philox_exp20(const Philox& ph, uint32_t r, uint32_t t, float* q) {
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    const uint4 x = ph(r, t, RS_SEQ, c);
    q[c * 4 + 0] = -logf(u01_open(x.x)); q[c * 4 + 1] = -logf(u01_open(x.y));
    q[c * 4 + 2] = -logf(u01_open(x.z)); q[c * 4 + 3] = -logf(u01_open(x.w));
  }
}

// ---------------------------------------------------------------- fused reverse step (dpm_full.py:284-298)

__global__ void __launch_bounds__(128)
step_kernel(StepArgs a, DiffW dw) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.M) return;
  const int t = a.t;
  const bool gen = a.mask_gen[r] != 0;
  const bool parity = a.nz.u != nullptr;
  const bool noisy = t > 1;
  Philox ph(a.seed);
  const uint32_t gr = (uint32_t)r + a.row0;      // Philox counter = row in the GLOBAL batch: a sharded run draws what the unsharded one draws

  // normalise positions exactly as the loop does on re-reading traj[t] (dpm_full.py:276,148-150)
  float pt[3], vt[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    pt[i] = div_(add_(a.p_t_ang[(size_t)r * 3 + i], -dw.pos_mean[i]), dw.pos_scale);
    vt[i] = a.v_t[(size_t)r * 3 + i];
  }
  const long long st = a.s_t[r];

  // ---- rotation (transition.py:146-160)
  float vo[3] = {vt[0], vt[1], vt[2]};
  if (gen && a.sample_structure) {
    float u[3], unif, gauss, ucdf = 0.f;
    if (parity) {
      u[0] = a.nz.u[(size_t)r * 3]; u[1] = a.nz.u[(size_t)r * 3 + 1]; u[2] = a.nz.u[(size_t)r * 3 + 2];
      unif = a.nz.unif_ang[r]; gauss = a.nz.gauss_ang[r];
    } else {
      const uint4 x = ph(gr, t, RS_U, 0);
      float g3;
      box_muller(x.x, x.y, u[0], u[1]); box_muller(x.z, x.w, u[2], g3);
      const uint4 y = ph(gr, t, RS_ANGLE, 0);
      float g1;
      ucdf = u01_half(y.x); unif = u01_half(y.y); box_muller(y.z, y.w, gauss, g1);
    }
    const float theta = sample_angle(dw, 1, t, parity ? a.nz.bin_idx[r] : 0, ucdf, unif, gauss, parity);
    compose_noise_rotation(u[0], u[1], u[2], theta, noisy, a.v_net[(size_t)r * 3], a.v_net[(size_t)r * 3 + 1],
                           a.v_net[(size_t)r * 3 + 2], vo[0], vo[1], vo[2]);
  }
  // ---- position (transition.py:42-50, 80-101)
  float po[3] = {pt[0], pt[1], pt[2]};
  if (gen && a.sample_structure) {
    float z[3];
    if (parity) { z[0] = a.nz.z_pos[(size_t)r * 3]; z[1] = a.nz.z_pos[(size_t)r * 3 + 1]; z[2] = a.nz.z_pos[(size_t)r * 3 + 2]; }
    else { const uint4 x = ph(gr, t, RS_ZPOS, 0); float g3; box_muller(x.x, x.y, z[0], z[1]); box_muller(x.z, x.w, z[2], g3); }
    const float alpha = fmaxf(dw.alphas[t], dw.alphas[dw.num_steps - 1]);   // clamp_min(alphas[-2])
    const float alpha_bar = dw.alpha_bars[t], sigma = dw.sigmas[t];
    const float c0 = div_(1.0f, sqrtf(add_(alpha, 1e-8f)));
    const float c1 = div_(add_(1.f, -alpha), sqrtf(add_(add_(1.f, -alpha_bar), 1e-8f)));
    const float ra = dw.sqrt_recip_ab[t], rb = dw.sqrt_recipm1_ab[t];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float pp = a.p_pred[(size_t)r * 3 + i];
      const float eps = a.pred_x0 ? div_(add_(mul_(ra, pt[i]), -pp), rb) : pp;
      const float zi = noisy ? z[i] : 0.f;
      po[i] = add_(mul_(c0, add_(pt[i], -mul_(c1, eps))), mul_(sigma, zi));
    }
  }
  // ---- sequence: the reference samples EVERY row, context and padding included (transition.py:241-244)
  float q[NAA];
  if (parity) {
#pragma unroll
    for (int k = 0; k < NAA; ++k) q[k] = a.nz.expo_seq[(size_t)r * NAA + k];
  } else {
    philox_exp20(ph, gr, t, q);
  }
  float c0v[NAA];
#pragma unroll
  for (int k = 0; k < NAA; ++k) c0v[k] = a.c_den[(size_t)r * NAA + k];
  float maxprob;
  const int sn = seq_posterior_sample(st, c0v, dw.alpha_bars_seq[t], gen, q, nullptr, maxprob);

#pragma unroll
  for (int i = 0; i < 3; ++i) {
    a.v_out[(size_t)r * 3 + i] = vo[i];
    a.p_out_ang[(size_t)r * 3 + i] = add_(mul_(po[i], dw.pos_scale), dw.pos_mean[i]);   // _unnormalize_position
  }
  a.s_out[r] = a.sample_sequence ? (long long)sn : st;
  if (a.maxprob_rows) a.maxprob_rows[r] = (gen || !a.masked_ppl) ? maxprob : 0.f;
}

// per-complex pRMSD score and perplexity (prmsd.py:31-47, dpm_full.py:380-399); one warp per complex
__global__ void complex_reduce_kernel(int N, int L, int bins, float dmin, float dmax, int masked_ppl,
                                      const float* __restrict__ prmsd_logits, const float* __restrict__ maxprob_rows,
                                      const uint8_t* __restrict__ mask_gen, float* __restrict__ prmsd_out,
                                      float* __restrict__ ppl_out) {
  const int n = blockIdx.x, lane = threadIdx.x;
  float s = 0.f, cnt = 0.f;
  for (int l = lane; l < L; l += 32) {
    s += maxprob_rows[(size_t)n * L + l];
    cnt += (!masked_ppl || mask_gen[(size_t)n * L + l]) ? 1.f : 0.f;
  }
  s = warp_sum(s); cnt = warp_sum(cnt);
  if (lane == 0) ppl_out[n] = s / cnt;
  if (prmsd_logits != nullptr) {
    float lg[2], mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < 2; ++u) { const int k = lane + u * 32; lg[u] = (k < bins) ? prmsd_logits[(size_t)n * bins + k] : -INFINITY; mx = fmaxf(mx, lg[u]); }
    mx = warp_max(mx);
    float se = 0.f, sw = 0.f;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int k = lane + u * 32;
      if (k < bins) {
        const float e = expf(lg[u] - mx);
        const float bound = dmin + (dmax - dmin) * (float)k / (float)(bins - 1);      // torch.linspace
        se += e; sw += e * bound;
      }
    }
    se = warp_sum(se); sw = warp_sum(sw);
    if (lane == 0) prmsd_out[n] = sw / se;
  }
}

// ---------------------------------------------------------------- DiffusionAntibodyDesign.encode, the per-residue part
// models/diffab.py:46-50,78-84,138: context_mask = mask_heavyatom[:, :, CA] & ~generate_flag; R_0 = construct_3d_basis(CA, C, N)
// (modules/common/geometry.py:47-69), v_0 = rotation_to_so3vec(R_0) (so3.py:10-30,60-63), p_0 = CA.  One thread per residue.
__global__ void __launch_bounds__(128)
design_prep_kernel(int M, int A_in, const float* __restrict__ pos, const uint8_t* __restrict__ mask_atoms,
                   const uint8_t* __restrict__ gen, uint8_t* __restrict__ ctx, float* __restrict__ v0, float* __restrict__ p0) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  const float* P = pos + (size_t)r * A_in * 3;                       // N 0..2, CA 3..5, C 6..8 (BBHeavyAtom order)
  const float ca[3] = {P[3], P[4], P[5]};
  float e1[3], v2[3], e2[3], e3[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { e1[c] = P[6 + c] - ca[c]; v2[c] = P[c] - ca[c]; }
  const float n1 = sqrtf(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]) + 1e-6f;
#pragma unroll
  for (int c = 0; c < 3; ++c) e1[c] = e1[c] / n1;
  const float pr = e1[0] * v2[0] + e1[1] * v2[1] + e1[2] * v2[2];
#pragma unroll
  for (int c = 0; c < 3; ++c) e2[c] = v2[c] - pr * e1[c];
  const float n2 = sqrtf(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]) + 1e-6f;
#pragma unroll
  for (int c = 0; c < 3; ++c) e2[c] = e2[c] / n2;
  e3[0] = e1[1] * e2[2] - e1[2] * e2[1]; e3[1] = e1[2] * e2[0] - e1[0] * e2[2]; e3[2] = e1[0] * e2[1] - e1[1] * e2[0];
  const Mat3 R = {{e1[0], e2[0], e3[0], e1[1], e2[1], e3[1], e1[2], e2[2], e3[2]}};       // columns e1 e2 e3
  float x, y, z;
  so3_log(R, x, y, z);
  v0[(size_t)r * 3] = x; v0[(size_t)r * 3 + 1] = y; v0[(size_t)r * 3 + 2] = z;
  p0[(size_t)r * 3] = ca[0]; p0[(size_t)r * 3 + 1] = ca[1]; p0[(size_t)r * 3 + 2] = ca[2];
  ctx[r] = (mask_atoms[(size_t)r * A_in + 1] != 0 && gen[r] == 0) ? 1 : 0;
}
void launch_design_prep(int M, int A_in, const float* pos, const uint8_t* mask_atoms, const uint8_t* gen, uint8_t* ctx, float* v0,
                        float* p0, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  design_prep_kernel<<<(M + 127) / 128, 128, 0, st>>>(M, A_in, pos, mask_atoms, gen, ctx, v0, p0);
}

// ---------------------------------------------------------------- initial state

__global__ void __launch_bounds__(128)
init_kernel(InitArgs a, DiffW dw) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < a.M / a.L && a.has_prmsd) { a.prmsd_out[r] = 0.f; a.ppl_out[r] = 1.f; }     // zeros_like / ones_like, dpm_full.py:269
  if (r >= a.M) return;
  const bool gen = a.mask_gen[r] != 0;
  Philox ph(a.seed);
  const uint32_t gr = (uint32_t)r + a.row0;
  const uint32_t tt = 0x7fffffffu;                 // "step" id of the initialisation draws
  float v[3], p[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    v[i] = a.v[(size_t)r * 3 + i];
    p[i] = div_(add_(a.p_ang[(size_t)r * 3 + i], -dw.pos_mean[i]), dw.pos_scale);
  }
  long long s = a.s[r];
  if (!a.optimize) {                                 // FullDPM.sample, dpm_full.py:254-267
    if (gen && a.sample_structure) {
      float g[4], gp3[3];
      if (a.g4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) g[i] = a.g4[(size_t)r * 4 + i];
#pragma unroll
        for (int i = 0; i < 3; ++i) gp3[i] = a.gp[(size_t)r * 3 + i];
      } else {
        const uint4 x = ph(gr, tt, RS_INIT_G4, 0); box_muller(x.x, x.y, g[0], g[1]); box_muller(x.z, x.w, g[2], g[3]);
        const uint4 y = ph(gr, tt, RS_INIT_GP, 0); float g3; box_muller(y.x, y.y, gp3[0], gp3[1]); box_muller(y.z, y.w, gp3[2], g3);
      }
      const Mat3 Rq = quat_to_rot(g[0], g[1], g[2], g[3]);          // random_uniform_so3, so3.py:66-68
      so3_log(Rq, v[0], v[1], v[2]);
      p[0] = gp3[0]; p[1] = gp3[1]; p[2] = gp3[2];
    }
    if (gen && a.sample_sequence) {
      if (a.s_rand) s = a.s_rand[r];
      else { const uint4 x = ph(gr, tt, RS_INIT_S, 0); s = (long long)(x.x % 19u); }   // randint_like(s, 0, 19): class 19 never drawn
    }
  } else {                                           // FullDPM.optimize, dpm_full.py:321-337
    const int t = a.tvec ? (int)a.tvec[r / a.L] : a.T0;
    const bool parity = a.add.u != nullptr;
    if (gen && a.sample_structure) {
      float u[3], unif, gauss, ucdf = 0.f, z[3];
      if (parity) {
        u[0] = a.add.u[(size_t)r * 3]; u[1] = a.add.u[(size_t)r * 3 + 1]; u[2] = a.add.u[(size_t)r * 3 + 2];
        unif = a.add.unif_ang[r]; gauss = a.add.gauss_ang[r];
        z[0] = a.add.z_pos[(size_t)r * 3]; z[1] = a.add.z_pos[(size_t)r * 3 + 1]; z[2] = a.add.z_pos[(size_t)r * 3 + 2];
      } else {
        const uint4 x = ph(gr, tt, RS_U, 0); float g3; box_muller(x.x, x.y, u[0], u[1]); box_muller(x.z, x.w, u[2], g3);
        const uint4 y = ph(gr, tt, RS_ANGLE, 0); float g1; ucdf = u01_half(y.x); unif = u01_half(y.y); box_muller(y.z, y.w, gauss, g1);
        const uint4 w = ph(gr, tt, RS_ZPOS, 0); box_muller(w.x, w.y, z[0], z[1]); box_muller(w.z, w.w, z[2], g3);
      }
      const float abr = dw.alpha_bars_rot[t], ab = dw.alpha_bars[t];
      const float c0r = sqrtf(abr);
      const float c0 = sqrtf(ab), c1 = sqrtf(add_(1.f, -ab));
      const float theta = sample_angle(dw, 0, t, parity ? a.add.bin_idx[r] : 0, ucdf, unif, gauss, parity);
      float nv[3];                                   // transition.py:120-144: log( exp(e) exp(c0 v_0) )
      compose_noise_rotation(u[0], u[1], u[2], theta, true, mul_(c0r, v[0]), mul_(c0r, v[1]), mul_(c0r, v[2]), nv[0], nv[1], nv[2],
                             a.grad_clamp ? -0.999f : -1.f);
      v[0] = nv[0]; v[1] = nv[1]; v[2] = nv[2];
#pragma unroll
      for (int i = 0; i < 3; ++i) p[i] = add_(mul_(c0, p[i]), mul_(c1, z[i]));     // transition.py:62-78
    }
    if (a.z_out && a.sample_structure) {               // e_rand is returned for every row (transition.py:74-78)
      float z[3];
      if (parity) { z[0] = a.add.z_pos[(size_t)r * 3]; z[1] = a.add.z_pos[(size_t)r * 3 + 1]; z[2] = a.add.z_pos[(size_t)r * 3 + 2]; }
      else { const uint4 w = ph(gr, tt, RS_ZPOS, 0); float g3; box_muller(w.x, w.y, z[0], z[1]); box_muller(w.z, w.w, z[2], g3); }
      a.z_out[(size_t)r * 3] = z[0]; a.z_out[(size_t)r * 3 + 1] = z[1]; a.z_out[(size_t)r * 3 + 2] = z[2];
    }
    if (a.sample_sequence) {                          // transition.py:183-200 on every row; kept only where generated
      float q[NAA];
      if (parity) {
#pragma unroll
        for (int k = 0; k < NAA; ++k) q[k] = a.add.expo_seq[(size_t)r * NAA + k];
      } else philox_exp20(ph, gr, tt, q);
      if (gen || a.seq_all_rows) {
        const float ab = gen ? dw.alpha_bars_seq[t] : 1.f;          // not generated: c_t = c_0 (transition.py:198)
        const float base = gen ? div_(add_(1.f, -ab), (float)NAA) : 0.f;
        const bool s_ok = (s >= 0 && s < NAA);
        float best = -INFINITY; int bi = 0;
#pragma unroll
        for (int k = 0; k < NAA; ++k) {
          const float c0k = (s_ok && k == (int)s) ? 1.f : 0.f;
          const float ck = add_(mul_(ab, c0k), base);
          const float vq = div_(add_(ck, 1e-8f), q[k]);
          if (vq > best) { best = vq; bi = k; }
        }
        s = bi;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    a.v_out[(size_t)r * 3 + i] = v[i];
    a.p_out_ang[(size_t)r * 3 + i] = add_(mul_(p[i], dw.pos_scale), dw.pos_mean[i]);
  }
  a.s_out[r] = s;
}

// ---------------------------------------------------------------- training losses (FullDPM.forward)
__global__ void gather_beta_kernel(int N, const long long* tvec, const float* betas, float* out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < N) out[n] = betas[tvec[n]];
}

// one thread per residue: the per-residue terms of the rot / pos / seq losses, the squared deviation behind the pRMSD
// target and one row of the distance loss.  dpm_full.py:186-232, :15-32, :369-378; transition.py:202-227
__global__ void __launch_bounds__(128)
loss_rows_kernel(LossArgs a, DiffW dw) {
  const int M = a.N * a.L;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  const int n = r / a.L;
  const int t = (int)a.tvec[n];
  const bool gen = a.mask_gen[r] != 0;
  float rot = 0.f, pos = 0.f, seq = 0.f, sq = 0.f, dsum = 0.f, dcnt = 0.f;
  float p0[3], pn[3], pp[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    p0[i] = div_(add_(a.p_0_ang[(size_t)r * 3 + i], -dw.pos_mean[i]), dw.pos_scale);
    pn[i] = div_(add_(a.p_noisy_ang[(size_t)r * 3 + i], -dw.pos_mean[i]), dw.pos_scale);
    pp[i] = a.p_pred[(size_t)r * 3 + i];
  }
  if (gen) {
    // ---- rotation: sum over columns of 1 - cos(col_pred, col_true)   (cosine_embedding_loss, EPSILON = 1e-12)
    const Mat3 R0 = so3_exp(a.v_0[(size_t)r * 3], a.v_0[(size_t)r * 3 + 1], a.v_0[(size_t)r * 3 + 2]);
    const float* Rp = a.R_pred + (size_t)r * 9;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x0 = Rp[c], x1 = Rp[3 + c], x2 = Rp[6 + c];
      const float y0 = R0.m[c], y1 = R0.m[3 + c], y2 = R0.m[6 + c];
      const float dot = x0 * y0 + x1 * y1 + x2 * y2;
      const float m1 = x0 * x0 + x1 * x1 + x2 * x2 + 1e-12f, m2 = y0 * y0 + y1 * y1 + y2 * y2 + 1e-12f;
      rot += 1.f - dot / sqrtf(m1 * m2);
    }
    // ---- position: AbDesign compares the predicted noise with e_rand; AbDock p_pred with p_0 (pred_x0) or p_noisy (sic)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float tgt = a.abdock ? (a.pred_x0 ? p0[i] : pn[i]) : (a.z ? a.z[(size_t)r * 3 + i] : 0.f);
      const float d = pp[i] - tgt;
      pos += d * d;
    }
    // ---- sequence: KL( posterior(s_noisy, s_0) || posterior(s_noisy, c_denoised) )
    const float ab = dw.alpha_bars_seq[t];
    const float base = (1.f - ab) / (float)NAA;
    const long long st = a.s_noisy[r], s0 = a.s_0[r];
    float tt_[NAA], tp_[NAA], st_sum = 0.f, sp_sum = 0.f;
#pragma unroll
    for (int k = 0; k < NAA; ++k) {
      const float ct = (st == k) ? 1.f : 0.f, c0 = (s0 == k) ? 1.f : 0.f;
      const float lhs = ab * ct + base;
      tt_[k] = lhs * (ab * c0 + base);
      tp_[k] = lhs * (ab * a.c_den[(size_t)r * NAA + k] + base);
      st_sum += tt_[k]; sp_sum += tp_[k];
    }
#pragma unroll
    for (int k = 0; k < NAA; ++k) {
      const float tg = tt_[k] / (st_sum + 1e-8f);
      const float lp = logf(tp_[k] / (sp_sum + 1e-8f) + 1e-8f);
      if (tg > 0.f) seq += tg * (logf(tg) - lp);
    }
    // ---- pRMSD target: |pred_p0 - p_0|^2 in Angstrom over generated residues   (common/prmsd.py:86-111)
    if (a.abdock) {
      const float c0 = dw.sqrt_recip_ab[t], c1 = dw.sqrt_recipm1_ab[t];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float pred0 = a.pred_x0 ? pp[i] : (c0 * p0[i] - c1 * pp[i]);      // pred_start_from_noise(p_0, p_pred, ...)
        const float d = (pred0 * dw.pos_scale + dw.pos_mean[i]) - (p0[i] * dw.pos_scale + dw.pos_mean[i]);
        sq += d * d;
      }
    }
    // ---- distance loss row: SmoothL1(|p_pred_i - p_pred_j| - |p_0i - p_0j|) over valid residues j
    if (a.abdock && a.pred_x0 && a.mask_res[r]) {
      const int base_r = n * a.L;
      for (int j = 0; j < a.L; ++j) {
        if (!a.mask_res[base_r + j]) continue;
        float dp = 0.f, dt = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float qp = a.p_pred[(size_t)(base_r + j) * 3 + i];
          const float q0 = div_(add_(a.p_0_ang[(size_t)(base_r + j) * 3 + i], -dw.pos_mean[i]), dw.pos_scale);
          dp += (pp[i] - qp) * (pp[i] - qp);
          dt += (p0[i] - q0) * (p0[i] - q0);
        }
        const float x = fabsf(sqrtf(dp) - sqrtf(dt));
        dsum += x < 1.f ? 0.5f * x * x : x - 0.5f;
        dcnt += 1.f;
      }
    }
  }
  a.rows[r] = rot; a.rows[(size_t)M + r] = pos; a.rows[(size_t)2 * M + r] = seq; a.rows[(size_t)3 * M + r] = sq;
  a.rows[(size_t)4 * M + r] = dsum; a.rows[(size_t)5 * M + r] = dcnt;
}

__device__ double block_sum_double(double v, double* sh) {
  const int tid = threadIdx.x;
  __syncthreads();
  sh[tid] = v;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) { if (tid < s) sh[tid] += sh[tid + s]; __syncthreads(); }
  return sh[0];
}

// one block: masked means in a fixed order (deterministic), pRMSD cross entropy per complex
__global__ void __launch_bounds__(1024)
loss_reduce_kernel(LossArgs a) {
  __shared__ double sh[1024];
  const int M = a.N * a.L, tid = threadIdx.x;
  double acc[6] = {0, 0, 0, 0, 0, 0}, ngen = 0;
  for (int r = tid; r < M; r += blockDim.x) {
#pragma unroll
    for (int k = 0; k < 6; ++k) if (k != 3) acc[k] += (double)a.rows[(size_t)k * M + r];
    ngen += a.mask_gen[r] ? 1.0 : 0.0;
  }
  const double n_gen = block_sum_double(ngen, sh);
  const double rot = block_sum_double(acc[0], sh), pos = block_sum_double(acc[1], sh), seq = block_sum_double(acc[2], sh);
  const double dsum = block_sum_double(acc[4], sh), dcnt = block_sum_double(acc[5], sh);
  // pRMSD: rmsd per complex -> nearest bin of linspace(dmin, dmax, bins) -> cross entropy; averaged with mask_generate[:, 0]
  double ce_sum = 0, m_sum = 0;
  if (a.abdock && a.has_prmsd) {
    for (int n = tid; n < a.N; n += blockDim.x) {
      float s2 = 0.f, cnt = 0.f;
      for (int l = 0; l < a.L; ++l) { s2 += a.rows[(size_t)3 * M + (size_t)n * a.L + l]; cnt += a.mask_gen[(size_t)n * a.L + l] ? 1.f : 0.f; }
      const float rmsd = sqrtf(s2 / cnt);
      const float step = (a.dmax - a.dmin) / (float)(a.bins - 1);
      int best = 0; float bd = INFINITY;
      for (int k = 0; k < a.bins; ++k) {
        const float off = (k < a.bins / 2) ? a.dmin + step * (float)k : a.dmax - step * (float)(a.bins - 1 - k);     // torch.linspace
        const float d = fabsf(rmsd - off);
        if (d < bd) { bd = d; best = k; }
      }
      const float* lg = a.prmsd_logits + (size_t)n * a.bins;
      float mx = -INFINITY;
      for (int k = 0; k < a.bins; ++k) mx = fmaxf(mx, lg[k]);
      float se = 0.f;
      for (int k = 0; k < a.bins; ++k) se += expf(lg[k] - mx);
      const float ce = (mx + logf(se)) - lg[best];
      const float mk = a.mask_gen[(size_t)n * a.L] ? 1.f : 0.f;
      ce_sum += (double)(ce * mk); m_sum += (double)mk;
    }
  }
  const double ce = block_sum_double(ce_sum, sh), ms = block_sum_double(m_sum, sh);
  if (tid == 0) {
    const double den = n_gen + 1e-8;
    a.out[0] = (float)(rot / den); a.out[1] = (float)(pos / den); a.out[2] = (float)(seq / den);
    a.out[3] = (a.abdock && a.has_prmsd) ? (float)(ce / (ms + 1e-10)) : 0.f;
    a.out[4] = (a.abdock && a.pred_x0) ? (float)(dsum / dcnt) : 0.f;
  }
}

// ---------------------------------------------------------------- stand-alone transition entry points
__global__ void rot_denoise_kernel(int M, int L, const float* v_t, const float* v_net, const uint8_t* mask_gen,
                                   const long long* tvec, NoisePtrs nz, DiffW dw, float* v_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  const int t = (int)tvec[r / L];
  float vo[3] = {v_t[(size_t)r * 3], v_t[(size_t)r * 3 + 1], v_t[(size_t)r * 3 + 2]};
  if (mask_gen[r]) {
    const float theta = sample_angle(dw, 1, t, nz.bin_idx[r], 0.f, nz.unif_ang[r], nz.gauss_ang[r], true);
    compose_noise_rotation(nz.u[(size_t)r * 3], nz.u[(size_t)r * 3 + 1], nz.u[(size_t)r * 3 + 2], theta, t > 1,
                           v_net[(size_t)r * 3], v_net[(size_t)r * 3 + 1], v_net[(size_t)r * 3 + 2], vo[0], vo[1], vo[2]);
  }
  v_out[(size_t)r * 3] = vo[0]; v_out[(size_t)r * 3 + 1] = vo[1]; v_out[(size_t)r * 3 + 2] = vo[2];
}

__global__ void pos_kernel(int M, int L, int mode, const float* p_t, const float* other, const uint8_t* mask_gen,
                           const long long* tvec, const float* z_pos, DiffW dw, float* out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  const int t = (int)tvec[r / L];
  const bool gen = mask_gen[r] != 0;
  for (int i = 0; i < 3; ++i) {
    const float pt = p_t[(size_t)r * 3 + i], ot = other[(size_t)r * 3 + i];
    float o = pt;
    if (gen) {
      if (mode == 0) {                  // pred_noise_from_start(p_t, p_0)
        o = div_(add_(mul_(dw.sqrt_recip_ab[t], pt), -ot), dw.sqrt_recipm1_ab[t]);
      } else {                          // denoise(p_t, eps_p)
        const float alpha = fmaxf(dw.alphas[t], dw.alphas[dw.num_steps - 1]);
        const float c0 = div_(1.0f, sqrtf(add_(alpha, 1e-8f)));
        const float c1 = div_(add_(1.f, -alpha), sqrtf(add_(add_(1.f, -dw.alpha_bars[t]), 1e-8f)));
        const float zi = (t > 1) ? z_pos[(size_t)r * 3 + i] : 0.f;
        o = add_(mul_(c0, add_(pt, -mul_(c1, ot))), mul_(dw.sigmas[t], zi));
      }
    }
    out[(size_t)r * 3 + i] = o;
  }
}

__global__ void seq_denoise_kernel(int M, int L, const long long* s_t, const float* c0, const uint8_t* mask_gen,
                                   const long long* tvec, const float* expo_seq, DiffW dw, float* post, long long* s_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  const int t = (int)tvec[r / L];
  float c0v[NAA], q[NAA], pst[NAA], mp;
  for (int k = 0; k < NAA; ++k) { c0v[k] = c0[(size_t)r * NAA + k]; q[k] = expo_seq[(size_t)r * NAA + k]; }
  const int sn = seq_posterior_sample(s_t[r], c0v, dw.alpha_bars_seq[t], mask_gen[r] != 0, q, pst, mp);
  for (int k = 0; k < NAA; ++k) post[(size_t)r * NAA + k] = pst[k];
  s_out[r] = sn;
}

// ---------------------------------------------------------------- launchers
void launch_angle_argmax(int M, int L, const long long* tvec, int t_uniform, const float* Y, const float* expo,
                         const uint8_t* mask_gen, int* bin_idx, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  angle_argmax_kernel<<<M, 256, 0, st>>>(M, L, tvec, t_uniform, Y, expo, mask_gen, bin_idx);
}
void launch_step(const StepArgs& a, const DiffW& dw, cudaStream_t st) {
  ProfScope prof__(KK_STEP, st);
  step_kernel<<<(a.M + 127) / 128, 128, 0, st>>>(a, dw);
}
void launch_complex_reduce(int N, int L, int bins, float dmin, float dmax, int masked_ppl, const float* prmsd_logits,
                           const float* maxprob_rows, const uint8_t* mask_gen, float* prmsd_out, float* ppl_out, cudaStream_t st) {
  ProfScope prof__(KK_STEP, st);
  complex_reduce_kernel<<<N, 32, 0, st>>>(N, L, bins, dmin, dmax, masked_ppl, prmsd_logits, maxprob_rows, mask_gen, prmsd_out, ppl_out);
}
void launch_init(const InitArgs& a, const DiffW& dw, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  init_kernel<<<(a.M + 127) / 128, 128, 0, st>>>(a, dw);
}
void launch_loss(const LossArgs& a, const DiffW& dw, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  const int M = a.N * a.L;
  loss_rows_kernel<<<(M + 127) / 128, 128, 0, st>>>(a, dw);
  loss_reduce_kernel<<<1, 1024, 0, st>>>(a);
}
void launch_gather_beta(int N, const long long* tvec, const float* betas, float* out, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  gather_beta_kernel<<<(N + 127) / 128, 128, 0, st>>>(N, tvec, betas, out);
}
void launch_rot_denoise(int M, int L, const float* v_t, const float* v_net, const uint8_t* mask_gen, const long long* tvec,
                        const NoisePtrs& nz, const DiffW& dw, float* v_out, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  rot_denoise_kernel<<<(M + 127) / 128, 128, 0, st>>>(M, L, v_t, v_net, mask_gen, tvec, nz, dw, v_out);
}
void launch_pos(int M, int L, int mode, const float* p_t, const float* other, const uint8_t* mask_gen, const long long* tvec,
                const float* z_pos, const DiffW& dw, float* out, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  pos_kernel<<<(M + 127) / 128, 128, 0, st>>>(M, L, mode, p_t, other, mask_gen, tvec, z_pos, dw, out);
}
void launch_seq_denoise(int M, int L, const long long* s_t, const float* c0, const uint8_t* mask_gen, const long long* tvec,
                        const float* expo_seq, const DiffW& dw, float* post, long long* s_out, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  seq_denoise_kernel<<<(M + 127) / 128, 128, 0, st>>>(M, L, s_t, c0, mask_gen, tvec, expo_seq, dw, post, s_out);
}

}  // namespace abopt
