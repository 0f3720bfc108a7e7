/*
 * abopt_b200.h -- C ABI of libabopt_b200.so: the B200 (sm_100a) implementation of ab_opt's
 * reverse-diffusion sampling hot path.
 *
 * The reference (pengzhangzhi/ab_opt) has no FFI of its own: its seam is Python class
 * substitution behind the model registry.  Each entry point below
 * therefore replaces one reference *method*; the citation gives the method it stands in for
 * (paths relative to the reference root, AbDock flavour; the AbDesign mirror is noted where it
 * differs).  INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - plain C types only; tensors are raw pointers + sizes, row-major contiguous, with the
 *     reference's shapes: N complexes, L residues, F=128 node channels, C=64 pair channels.
 *   - fp32 for every floating tensor, int64 for amino-acid indices and step indices, one byte
 *     per element (0/1) for masks -- exactly torch.float32 / torch.int64 / torch.bool storage.
 *   - "dev" pointers are device memory on the model's device; "host" pointers are host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Device entry
 *     points only enqueue work; they never synchronise.  Host entry points return after the
 *     outputs are complete in host memory.
 *   - every function returns ABOPT_OK (0) or a negative error code; abopt_last_error() gives a
 *     thread-local message.  Nothing falls back to the CPU: if no CUDA device of compute
 *     capability 10.x is usable the calls fail.
 */
#ifndef ABOPT_B200_H_
#define ABOPT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABOPT_OK              0
#define ABOPT_ERR_ARG        -1   /* bad argument / unsupported shape                       */
#define ABOPT_ERR_CUDA       -2   /* CUDA runtime error (message has the cudaError string)   */
#define ABOPT_ERR_STATE      -3   /* model not finalised, tensor missing, ...               */
#define ABOPT_ERR_KEY        -4   /* unknown state-dict key or wrong element count          */

#define ABOPT_NODE_DIM      128   /* res_feat_dim  (configs/train/dock_single.yml:3)        */
#define ABOPT_PAIR_DIM       64   /* pair_feat_dim (configs/train/dock_single.yml:4)        */
#define ABOPT_NUM_HEADS      12   /* modules/encoders/ga.py:43                              */
#define ABOPT_NUM_AA         20   /* modules/diffusion/transition.py:165                    */
#define ABOPT_ANGLE_BINS   8192   /* modules/common/so3.py:73                               */
#define ABOPT_MAX_L         512   /* longest complex: one attention row = 512 TMEM columns    */

typedef struct abopt_model abopt_model;

/* Hyper-parameters of FullDPM.__init__ (modules/diffusion/dpm_full.py:117-146). */
typedef struct abopt_config {
  int32_t num_layers;        /* eps_net_opt.num_layers (6)                                   */
  int32_t num_steps;         /* num_steps (100)                                              */
  int32_t has_prmsd;         /* 1 = AbDock flavour (pRMSD head, 5 trajectory fields)         */
  int32_t prmsd_bins;        /* num_bins (40); ignored when has_prmsd == 0                   */
  float   prmsd_min;         /* dist_min (0.5)                                               */
  float   prmsd_max;         /* dist_max (19.5)                                              */
  int32_t obj_pred_x0;       /* 1 = obj 'pred_x0', 0 = 'pred_noise' (dpm_full.py:143)        */
  int32_t scope;             /* ABOPT_SCOPE_*: which module this handle backs                */
} abopt_config;

/* A handle can back a whole FullDPM, or a stand-alone GAEncoder / EpsilonNet module: the scope
 * decides which state-dict keys are expected (always spelled as inside FullDPM, i.e. with the
 * "eps_net.encoder.blocks.N." / "eps_net." prefixes) and which entry points are usable. */
#define ABOPT_SCOPE_FULL     0   /* FullDPM: every entry point                                */
#define ABOPT_SCOPE_ENCODER  1   /* GAEncoder: abopt_ga_* only                                */
#define ABOPT_SCOPE_EPSNET   2   /* EpsilonNet: abopt_ga_* and abopt_eps_net_forward          */

/* Flags of abopt_sample_*. */
#define ABOPT_SAMPLE_STRUCTURE   1u   /* sample_structure=True                               */
#define ABOPT_SAMPLE_SEQUENCE    2u   /* sample_sequence=True                                */
#define ABOPT_KEEP_TRAJECTORY    4u   /* fill every trajectory slot (else only slot 0 and T0) */
#define ABOPT_GRAD_SEMANTICS     8u   /* abopt_loss_forward: evaluate as with autograd enabled (log_rotation clamp -0.999) */

/* Per-step noise for replayed ("parity") runs; every pointer is a DEVICE pointer holding the
 * draw the reference makes at that point (RNG draw order):
 *   u         (N,L,3)      randn            modules/common/so3.py:143
 *   expo_ang  (N*L,8191)   exponential_(1)  inside multinomial, so3.py:123
 *   unif_ang  (N*L)        rand             so3.py:126
 *   gauss_ang (N*L)        randn            so3.py:131
 *   z_pos     (N,L,3)      randn            modules/diffusion/transition.py:95
 *   expo_seq  (N*L,20)     exponential_(1)  inside multinomial, transition.py:179          */
typedef struct abopt_step_noise {
  const float* u;
  const float* expo_ang;
  const float* unif_ang;
  const float* gauss_ang;
  const float* z_pos;
  const float* expo_seq;
} abopt_step_noise;

/* ------------------------------------------------------------------ library / model life cycle */
int         abopt_version(void);
const char* abopt_last_error(void);
/* Number of kernels this library has launched in the calling process (all models). */
uint64_t    abopt_kernel_launch_count(void);

/* Optional per-kernel timing (bench.py's roofline line): while enabled, every kernel launch of this
 * library is bracketed by a CUDA event pair on its stream.  abopt_profile_collect() synchronises,
 * sums milliseconds and launch counts per kernel kind and clears the records.  Kinds, in order:
 * 0 mixer, 1 projections, 2 logits, 3 pair stream, 4 aggregation, 5 block tail, 6 heads,
 * 7 transition step, 8 other, 9 context-cache delta (first block inside the sampling loop), 10 pair stream launches that visit
 * the generated query rows only (3 = the full-stream launches).  Not for use inside timed regions (events perturb the pipeline). */
#define ABOPT_KERNEL_KINDS 11
int abopt_profile_enable(int on);
int abopt_profile_collect(double* ms_per_kind, uint64_t* launches_per_kind, int n_kinds);

/* Test hook for the tcgen05 "3xTF32" GEMM building block used by the linear layers:
 * D[M][N] = A[M][K] * B[N][K]^T (+ bias[N]); device pointers; N % 4 == 0, K % 32 == 0.  Synchronises. */
int abopt_debug_gemm3x(int device, int M, int N, int K, const float* A, const float* B, const float* bias,
                       float* D, void* stream);

/* Debug hook: SM-clock timestamps of the phases of one CTA of the last attention-logits kernel (16 slots). */
int abopt_debug_clocks(long long* out16);

/* Debug hook (parity bisection): copy one internal workspace tensor of the last GABlock call into `dst` (device memory with room
 * for max_floats): which = 0 QA, 1 KB, 2 rq, 3 rk, 4 VT (packed attention operands), 5 pair bias (slot 0), 6 alpha, 7 aggregate.
 * *numel receives the number of floats copied.  Enqueues on `stream`. */
int abopt_debug_copy(abopt_model* m, int which, float* dst, size_t max_floats, size_t* numel, void* stream);

/* FullDPM.__init__ : allocate an empty model on CUDA device `device`. */
int  abopt_model_create(const abopt_config* cfg, int device, abopt_model** out);
void abopt_model_destroy(abopt_model* m);

/* nn.Module.load_state_dict (tools/runner/design_for_pdb.py:94): copy one tensor of the FullDPM
 * state-dict, addressed by its reference key (e.g. "eps_net.encoder.blocks.0.proj_query.weight",
 * "trans_rot.angular_distrib_inv.Y", "position_scale").  `numel` must match the reference shape.
 * `dtype`: 0 = float32, 1 = uint8/bool, 2 = int64.  `on_device` != 0 if `data` is device memory. */
int abopt_model_set_tensor(abopt_model* m, const char* key, const void* data, size_t numel,
                           int dtype, int on_device);
/* Verify that every tensor is present and build the packed device layout. */
int abopt_model_finalize(abopt_model* m);
/* Batch sharding (the reference has no counterpart -- it runs one process on one device): index of this
 * handle's first complex inside the global, unsharded batch.  The in-kernel Philox counters are keyed by the GLOBAL residue
 * row, so N ranks that each run their contiguous slice (same seed, offset = first complex of the slice) reproduce the
 * single-device run of the whole batch bit for bit.  Default 0. */
int abopt_model_set_batch_offset(abopt_model* m, int64_t first_complex);

/* ------------------------------------------------------------------ encoder (device pointers) */
/* GABlock.forward, modules/encoders/ga.py:149-178.
 *   R (N,L,3,3)  t (N,L,3)  x (N,L,128)  z (N,L,L,64)  mask (N,L) u8  ->  x_out (N,L,128)   */
int abopt_ga_block_forward(abopt_model* m, int layer, int N, int L, const float* R, const float* t,
                           const float* x, const float* z, const uint8_t* mask, float* x_out,
                           void* stream);
/* GAEncoder.forward, modules/encoders/ga.py:190-193 (all layers). */
int abopt_ga_encoder_forward(abopt_model* m, int N, int L, const float* R, const float* t,
                             const float* x, const float* z, const uint8_t* mask, float* x_out,
                             void* stream);
/* Debug/parity taps of one block: attention weights alpha (N,L,L,12) in the reference layout and
 * the concatenated aggregate (N,L,1824) that feeds out_transform (ga.py:166-174).  Either output
 * may be NULL. */
int abopt_ga_block_taps(abopt_model* m, int layer, int N, int L, const float* R, const float* t,
                        const float* x, const float* z, const uint8_t* mask, float* alpha,
                        float* feat, void* stream);

/* ------------------------------------------------------------------ denoiser network */
/* EpsilonNet.forward, modules/diffusion/dpm_full.py:70-112 (AbDesign: :62-102).
 *   v_t,p_t (N,L,3)  s_t (N,L) i64  res_feat (N,L,128)  pair_feat (N,L,L,64)  beta (N,)
 *   -> v_next (N,L,3)  R_next (N,L,3,3)  eps_pos (N,L,3)  c_denoised (N,L,20)
 *      prmsd_logits (N,prmsd_bins) (has_prmsd models; may be NULL otherwise)                  */
int abopt_eps_net_forward(abopt_model* m, int N, int L, const float* v_t, const float* p_t,
                          const int64_t* s_t, const float* res_feat, const float* pair_feat,
                          const float* beta, const uint8_t* mask_generate, const uint8_t* mask_res,
                          float* v_next, float* R_next, float* eps_pos, float* c_denoised,
                          float* prmsd_logits, void* stream);

/* ------------------------------------------------------------------ transitions (device pointers)
 * `t` is the (N,) int64 step tensor the reference passes.  Noise pointers follow abopt_step_noise. */
/* RotationTransition.denoise, modules/diffusion/transition.py:146-160 (+ so3.py:111-146). */
int abopt_rot_denoise(abopt_model* m, int N, int L, const float* v_t, const float* v_net,
                      const uint8_t* mask_generate, const int64_t* t, const float* u,
                      const float* expo_ang, const float* unif_ang, const float* gauss_ang,
                      float* v_out, void* stream);
/* PositionTransition.pred_noise_from_start, transition.py:42-50. */
int abopt_pos_pred_noise_from_start(abopt_model* m, int N, int L, const float* p_t, const float* p_0,
                                    const uint8_t* mask_generate, const int64_t* t, float* eps_out,
                                    void* stream);
/* PositionTransition.denoise, transition.py:80-101. */
int abopt_pos_denoise(abopt_model* m, int N, int L, const float* p_t, const float* eps_p,
                      const uint8_t* mask_generate, const int64_t* t, const float* z_pos,
                      float* p_out, void* stream);
/* AminoacidCategoricalTransition.denoise, transition.py:229-245: post (N,L,20), s_out (N,L) i64. */
int abopt_seq_denoise(abopt_model* m, int N, int L, const int64_t* s_t, const float* c0_pred,
                      const uint8_t* mask_generate, const int64_t* t, const float* expo_seq,
                      float* post, int64_t* s_out, void* stream);

/* ------------------------------------------------------------------ sampling loop
 * FullDPM.sample (dpm_full.py:236-302) when opt_step == 0, FullDPM.optimize (:304-367) when
 * opt_step > 0.  Trajectory outputs are indexed by step: slot k holds traj[k] of the reference,
 * k = 0..T0 with T0 = num_steps (sample) or opt_step (optimize):
 *   traj_v (T0+1,N,L,3)  traj_p (T0+1,N,L,3) in Angstrom  traj_s (T0+1,N,L) i64
 *   traj_prmsd, traj_ppl (T0+1,N)  (has_prmsd models; may be NULL otherwise)
 * Without ABOPT_KEEP_TRAJECTORY only slots 0 and T0 are guaranteed to be filled.
 * Randomness: `noise` == NULL -> counter-based Philox4x32-10 keyed by `seed` inside the kernels
 * ("fast" mode, distributionally identical to the reference).  Otherwise `noise` points to T0
 * per-step records, noise[T0 - t] for step t, plus `init_noise` = the initialisation draws
 * (sample: g4 (N,L,4), gp (N,L,3) floats and s_rand (N,L) i64; optimize: one abopt_step_noise)
 * ("parity" mode: bit-identical consumption of the reference's draws).                        */
typedef struct abopt_init_noise {
  const float*   g4;      /* sample(): randn (N,L,4), so3.py:67                                */
  const float*   gp;      /* sample(): randn_like(p), dpm_full.py:257                          */
  const int64_t* s_rand;  /* sample(): randint_like(s, 0, 19), dpm_full.py:264                 */
  const abopt_step_noise* add;   /* optimize(): draws of the three add_noise calls             */
} abopt_init_noise;

/* The two halves of the loop, for callers that drive it step by step (the parity mode of the
 * Python FullDPM draws each step's noise with the reference's own ATen calls and hands it in):
 *   abopt_sample_init  = dpm_full.py:254-269 (sample) / :321-339 (optimize): inputs -> traj[T0]
 *   abopt_reverse_step = one iteration of dpm_full.py:274-300: traj[t] -> traj[t-1]
 * p / p_t / p_out are in Angstrom, as stored in the reference's trajectory.  `noise` / `init_noise`
 * may be NULL (Philox). prmsd_out / ppl_out (N,) may be NULL. */
int abopt_sample_init(abopt_model* m, int N, int L, const float* v, const float* p, const int64_t* s,
                      const uint8_t* mask_generate, uint32_t flags, int opt_step, uint64_t seed,
                      const abopt_init_noise* init_noise, float* v_out, float* p_out, int64_t* s_out,
                      void* stream);
int abopt_reverse_step(abopt_model* m, int N, int L, int t, int optimize, const float* v_t,
                       const float* p_t, const int64_t* s_t, const float* res_feat,
                       const float* pair_feat, const uint8_t* mask_generate, const uint8_t* mask_res,
                       uint32_t flags, uint64_t seed, const abopt_step_noise* noise, float* v_out,
                       float* p_out, int64_t* s_out, float* prmsd_out, float* ppl_out, void* stream);

int abopt_sample_device(abopt_model* m, int N, int L, const float* v, const float* p, const int64_t* s,
                        const float* res_feat, const float* pair_feat, const uint8_t* mask_generate,
                        const uint8_t* mask_res, uint32_t flags, int opt_step, uint64_t seed,
                        const abopt_init_noise* init_noise, const abopt_step_noise* noise,
                        float* traj_v, float* traj_p, int64_t* traj_s, float* traj_prmsd,
                        float* traj_ppl, void* stream);

/* Same, HOST buffers in and out: copies inputs to the device, runs the loop, copies the requested
 * trajectory slots back, synchronises.  This is the call design_pdb.py / dock_pdb.py would make
 * per batch; bench.py times it as the end-to-end number. */
int abopt_sample_host(abopt_model* m, int N, int L, const float* v, const float* p, const int64_t* s,
                      const float* res_feat, const float* pair_feat, const uint8_t* mask_generate,
                      const uint8_t* mask_res, uint32_t flags, int opt_step, uint64_t seed,
                      float* traj_v, float* traj_p, int64_t* traj_s, float* traj_prmsd,
                      float* traj_ppl);

/* ------------------------------------------------------------------ training forward
 * FullDPM.forward without autograd (dpm_full.py:156-234; AbDesign flavour :138-190): noise v_0 / p_0 / s_0 to the
 * per-complex steps t (N,) i64 with the three add_noise calls (transition.py:62-78,120-144,183-200), evaluate
 * EpsilonNet once and reduce the loss dict on the device.  flags: ABOPT_SAMPLE_STRUCTURE = denoise_structure,
 * ABOPT_SAMPLE_SEQUENCE = denoise_sequence.  p_0 in Angstrom.  `noise` == NULL -> Philox draws keyed by `seed`, else one
 * abopt_step_noise record = the reference's draws in its own order.  losses_out: DEVICE pointer to 5 floats
 *   [0] rot  [1] pos  [2] seq  [3] prmsd (has_prmsd models, else 0)  [4] dist (has_prmsd + obj pred_x0, else 0).
 * Evaluated as the reference evaluates it under torch.no_grad() (validate(), AbDock/train.py:141-149): log_rotation clamps the
 * cosine at -1.  With autograd enabled the reference clamps at -0.999 (modules/common/so3.py:12-17), which changes the noised
 * rotation of residues within 0.045 rad of pi: flag ABOPT_GRAD_SEMANTICS selects that evaluation here, and abopt_loss_backward
 * (below) always uses it. */
int abopt_loss_forward(abopt_model* m, int N, int L, const float* v_0, const float* p_0, const int64_t* s_0,
                       const float* res_feat, const float* pair_feat, const uint8_t* mask_generate,
                       const uint8_t* mask_res, uint32_t flags, const int64_t* t, uint64_t seed,
                       const abopt_step_noise* noise, float* losses_out, void* stream);

/* ------------------------------------------------------------------ training step: forward + backward
 * One iteration of AbDock/train.py:104-113 (AbDesign/train.py likewise) up to the optimiser: loss_dict = model(batch); loss =
 * sum_k w_k loss_k; loss.backward().  The reference differentiates FullDPM.forward (dpm_full.py:156-234) with torch autograd; here
 * the backward pass is written out (csrc/k_backward.cu), recompute-based: only the inputs of the GABlocks are kept.  The forward
 * is evaluated as the reference evaluates it WITH autograd enabled (log_rotation clamps the cosine at -0.999, so3.py:12-17), so
 * losses_out may differ from abopt_loss_forward (the torch.no_grad() evaluation) on rotations within 0.045 rad of pi.
 *   loss_weights  HOST pointer to 5 floats (rot, pos, seq, prmsd, dist; configs/train/*.yml loss_weights) or NULL (all 1)
 *   losses_out    DEVICE, 5 floats, the unweighted loss dict as abopt_loss_forward
 *   d_res_feat (N,L,128), d_pair_feat (N,L,L,64)   DEVICE, d loss / d inputs (they come from trainable embeddings)
 * Parameter gradients stay inside the handle (overwritten by every call); abopt_model_get_grad copies the one of a state-dict key
 * (same keys as abopt_model_set_tensor; buffers have none) to `dst` (DEVICE, `numel` floats).  Other arguments as abopt_loss_forward.
 * abopt_ga_block_backward is the same for one GABlock (ga.py:149-178): g_out = d loss / d block output -> g_x, g_z (+ the block's
 * parameter gradients in the handle). */
int abopt_loss_backward(abopt_model* m, int N, int L, const float* v_0, const float* p_0, const int64_t* s_0,
                        const float* res_feat, const float* pair_feat, const uint8_t* mask_generate, const uint8_t* mask_res,
                        uint32_t flags, const int64_t* t, uint64_t seed, const abopt_step_noise* noise, const float* loss_weights,
                        float* losses_out, float* d_res_feat, float* d_pair_feat, void* stream);
int abopt_model_get_grad(abopt_model* m, const char* key, float* dst, size_t numel, void* stream);
int abopt_ga_block_backward(abopt_model* m, int layer, int N, int L, const float* R, const float* t, const float* x,
                            const float* z, const uint8_t* mask, const float* g_out, float* g_x, float* g_z, void* stream);

/* ------------------------------------------------------------------ pair featurisation (the step before the loop)
 * PairEmbedding, modules/encoders/pair.py:10-101 (AbDesign: diffab/modules/encoders/pair.py, same lines), including
 * pairwise_dihedrals (modules/common/geometry.py:351-376) and AngularEncoding (modules/common/layers.py:85-106): builds
 * pair_feat (N,L,L,64), the loop-invariant `z` every GABlock streams.  Its own handle because it is a separate module of
 * DiffusionAntibodyDesign (models/diffab.py:28, state-dict prefix "pair_embed.").
 *   create      PairEmbedding.__init__(feat_dim=64, max_num_atoms) with max_aa_types=22, max_relpos=32 (pair.py:12);
 *               max_num_atoms = 15 ('full'), 5 ('backbone+CB') or 4 ('backbone') (models/diffab.py:13-17)
 *   set_tensor  one float32 tensor of PairEmbedding.state_dict(), by key: aa_pair_embed.weight (484,64), relpos_embed.weight
 *               (65,64), aapair_to_distcoef.weight (484,A*A), distance_embed.{0,2}.{weight,bias}, dihedral_embed.freq_bands (6),
 *               out_mlp.{0,2,4}.{weight,bias}; out_mlp.0.weight is (64, 218)
 *   forward     PairEmbedding.forward (pair.py:37-101).  DEVICE pointers: aa, res_nb, chain_nb (N,L) i64; pos_atoms
 *               (N,L,num_atoms_in,3) f32 in Angstrom and mask_atoms (N,L,num_atoms_in) u8 with num_atoms_in >= max_num_atoms
 *               (the first max_num_atoms atoms are used, pair.py:54-55); structure_mask / sequence_mask (N,L) u8 or NULL;
 *               pair_feat (N,L,L,64) f32 out.  Enqueues one kernel on `stream`, never synchronises.  Amino-acid indices
 *               outside [0, 22) are clamped (the reference's embedding lookup would raise). */
typedef struct abopt_pair_embed abopt_pair_embed;
int  abopt_pair_embed_create(int max_num_atoms, int device, abopt_pair_embed** out);
void abopt_pair_embed_destroy(abopt_pair_embed* pe);
int  abopt_pair_embed_set_tensor(abopt_pair_embed* pe, const char* key, const float* data, size_t numel, int on_device);
int  abopt_pair_embed_finalize(abopt_pair_embed* pe);
int  abopt_pair_embed_forward(abopt_pair_embed* pe, int N, int L, int num_atoms_in, const int64_t* aa, const int64_t* res_nb,
                              const int64_t* chain_nb, const float* pos_atoms, const uint8_t* mask_atoms,
                              const uint8_t* structure_mask, const uint8_t* sequence_mask, float* pair_feat, void* stream);

/* ResidueEmbedding, modules/encoders/residue.py:9-94 (state-dict prefix "residue_embed.", models/diffab.py:27): builds res_feat
 * (N,L,128) from the amino-acid type, the atom coordinates in the residue's own backbone frame (construct_3d_basis /
 * global_to_local, modules/common/geometry.py:47-69,94-113), the backbone dihedrals (geometry.py:307-348, topology.py:5-24)
 * and the fragment type.  Same life cycle as abopt_pair_embed_*.
 *   set_tensor  keys of ResidueEmbedding.state_dict(): aatype_embed.weight (22,128), type_embed.weight (10,128),
 *               dihed_embed.freq_bands (6), mlp.0.weight (256, 128 + 22*A*3 + 39 + 128), mlp.{2,4,6}.weight, mlp.{0,2,4,6}.bias
 *   forward     ResidueEmbedding.forward (residue.py:27-94); arguments as abopt_pair_embed_forward plus fragment_type (N,L) i64
 *               (clamped to [0, 10)); res_feat (N,L,128) f32 out. */
typedef struct abopt_res_embed abopt_res_embed;
int  abopt_res_embed_create(int max_num_atoms, int device, abopt_res_embed** out);
void abopt_res_embed_destroy(abopt_res_embed* re);
int  abopt_res_embed_set_tensor(abopt_res_embed* re, const char* key, const float* data, size_t numel, int on_device);
int  abopt_res_embed_finalize(abopt_res_embed* re);
int  abopt_res_embed_forward(abopt_res_embed* re, int N, int L, int num_atoms_in, const int64_t* aa, const int64_t* res_nb,
                             const int64_t* chain_nb, const float* pos_atoms, const uint8_t* mask_atoms,
                             const int64_t* fragment_type, const uint8_t* structure_mask, const uint8_t* sequence_mask,
                             float* res_feat, void* stream);

/* ------------------------------------------------------------------ after the loop (stateless, DEVICE pointers, current device)
 * reconstruct_backbone_partially, modules/common/geometry.py:450-480 (reconstruct_backbone :404-447): residues flagged in
 * mask_recons get ideal N, CA, C placed by their frame (R_new, t_new) and O placed after the psi turn, all other atoms zeroed
 * and masked out; the other residues pass through.  N may be frames x complexes (a whole trajectory in one launch).
 *   pos_ctx (N,L,A,3)  R_new (N,L,3,3)  t_new (N,L,3)  aa, chain_nb, res_nb (N,L) i64  mask_atoms (N,L,A) u8  mask_recons (N,L) u8
 *   bb_table (21,3,3), o_table (21,3): backbone_atom_coordinates_tensor / bb_oxygen_coordinate_tensor (utils/protein/constants.py:310-320)
 *   -> pos_new (N,L,A,3)  mask_new (N,L,A) u8 */
int abopt_reconstruct_backbone_partially(int N, int L, int A, const float* pos_ctx, const float* R_new, const float* t_new,
                                         const int64_t* aa, const int64_t* chain_nb, const int64_t* res_nb,
                                         const uint8_t* mask_atoms, const uint8_t* mask_recons, const float* bb_table,
                                         const float* o_table, float* pos_new, uint8_t* mask_new, void* stream);
/* calc_per_rmsd / calc_avg_rmsd, tools/runner/design_for_testset.py:556-570: structures (B,M,3) -> rmsd (B,B) (may be NULL),
 * score (B,) = row sums / (B-1) (the quantity rank_commoness sorts), avg (1,) = rmsd.sum() / (B (B-1)) (may be NULL). */
int abopt_pairwise_rmsd(int B, int M, const float* structures, float* rmsd, float* score, float* avg, void* stream);
/* rank_commoness, design_for_testset.py:573-589: rank (k,) i64 = indices of the k smallest scores, best first (ties: lower index
 * first); score (B,) scratch/out. */
int abopt_rank_commoness(int B, int M, const float* structures, int k, float* score, int64_t* rank, void* stream);

/* ------------------------------------------------------------------ encode + sample: atoms in, structures out
 * DiffusionAntibodyDesign.sample (models/diffab.py:115-141) / .optimize (:143-171; opt_step > 0): encode() -- context_mask =
 * mask_heavyatom[:, :, CA] & ~generate_flag (:46-50), ResidueEmbedding and PairEmbedding with structure_mask / sequence_mask =
 * context_mask where sample_structure / sample_sequence (:52-76,133-137), R_0 = construct_3d_basis(CA, C, N), p_0 = CA (:78-84),
 * v_0 = rotation_to_so3vec(R_0), s_0 = aa (:138-139) -- then FullDPM.sample / optimize, as ONE device-resident call: res_feat
 * and pair_feat (1.07 GB at N=64, L=256) are produced and consumed on the device.  Inputs are the reference's batch fields:
 *   aa, res_nb, chain_nb, fragment_type (N,L) i64; pos_heavyatom (N,L,num_atoms_in,3) f32 Angstrom; mask_heavyatom
 *   (N,L,num_atoms_in) u8; generate_flag, mask (N,L) u8.  pe / re: finalised embedding handles of the same device.
 * Trajectory outputs, flags, opt_step, seed as abopt_sample_device (Philox mode).  _device: DEVICE pointers, enqueues on `stream`;
 * _host: HOST pointers, the call design_pdb.py / dock_pdb.py make per batch (about 3 MB in; the trajectory out). */
int abopt_design_device(abopt_model* m, abopt_pair_embed* pe, abopt_res_embed* re, int N, int L, int num_atoms_in,
                        const int64_t* aa, const int64_t* res_nb, const int64_t* chain_nb, const float* pos_heavyatom,
                        const uint8_t* mask_heavyatom, const int64_t* fragment_type, const uint8_t* generate_flag,
                        const uint8_t* mask, uint32_t flags, int opt_step, uint64_t seed, float* traj_v, float* traj_p,
                        int64_t* traj_s, float* traj_prmsd, float* traj_ppl, void* stream);
int abopt_design_host(abopt_model* m, abopt_pair_embed* pe, abopt_res_embed* re, int N, int L, int num_atoms_in,
                      const int64_t* aa, const int64_t* res_nb, const int64_t* chain_nb, const float* pos_heavyatom,
                      const uint8_t* mask_heavyatom, const int64_t* fragment_type, const uint8_t* generate_flag,
                      const uint8_t* mask, uint32_t flags, int opt_step, uint64_t seed, float* traj_v, float* traj_p,
                      int64_t* traj_s, float* traj_prmsd, float* traj_ppl);

/* Size in bytes of the device scratch the model holds for (N, L); 0 if none allocated yet. */
size_t abopt_workspace_bytes(const abopt_model* m);

#ifdef __cplusplus
}
#endif
#endif /* ABOPT_B200_H_ */
