// C ABI of libabopt_b200 (see include/abopt_b200.h): model life cycle, weight packing, workspace and the
// orchestration of the kernels into GABlock / GAEncoder / EpsilonNet / FullDPM.sample|optimize.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "../../include/abopt_b200.h"
#include "kernels.h"

namespace abopt {
unsigned long long g_launches = 0;
// ---- optional profiler: one cudaEvent pair per launch, summed per kernel kind on collection
static bool g_prof_on = false;
struct ProfRec { int kind; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof;
static cudaEvent_t g_prof_open = nullptr;
void prof_begin(int kind, cudaStream_t st) {
  if (!g_prof_on) return;
  if (g_prof_open) cudaEventDestroy(g_prof_open);      // a begin without its end: drop the stale event
  cudaEventCreate(&g_prof_open);
  cudaEventRecord(g_prof_open, st);
}
void prof_end(int kind, cudaStream_t st) {
  if (!g_prof_on || !g_prof_open) return;
  cudaEvent_t b;
  cudaEventCreate(&b);
  cudaEventRecord(b, st);
  g_prof.push_back({kind, g_prof_open, b});
  g_prof_open = nullptr;
}
}
using namespace abopt;

// ------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
namespace abopt { int api_fail(int code, const std::string& msg) { return fail(code, msg); } }      // for the other translation units
#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return fail(ABOPT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));        \
  } while (0)
#define CHECK_LAUNCH() CUDA_TRY(cudaGetLastError())

struct DeviceGuard {
  int prev = -1; bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) { cudaSetDevice(dev); switched = true; }
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

// ------------------------------------------------------------------------------------------ model
struct HostTensor { std::vector<unsigned char> bytes; size_t numel = 0; int dtype = 0; bool set = false; };
struct KeySpec { size_t numel; int dtype; bool required; };

struct Workspace {
  int N = 0, L = 0, Lp = 0, NB = 0;
  size_t bytes = 0;
  void* base = nullptr;
  AttnOperands op;
  float *Rbuf, *pnorm, *xa, *xb, *xa_lo, *xb_lo, *xin_lo, *feat, *alpha, *prmsd_rows, *prmsd_logits, *maxprob;
  float *v_net, *eps_pos, *c_den, *R_next;
  int* bin_idx;
  long long* tvec_scratch;
  Focus focus;
  PairRows prows[3];            // row lists of pair_stream_kernel: [0] every row, [1] focus mode (generated rows, compact output),
                                // [2] generated rows in place (first block with the context cache)
  // loop invariants of the sampling loop (context residues): mixer output and the first block's packed operands keep their own
  // buffers, so that per step only the generated rows are recomputed (x0_c / x0_c_lo: those rows, compact, the projection's input)
  float *x0, *x0_lo, *x0_c, *x0_c_lo;
  AttnOperands op0;
  float* ctx_cache;             // [N * L][768]  context part of the first block's pair aggregate (k_pair.cu: ctx_delta_kernel)
  float2 *ctx_stats, *step_stats;      // [N][H][L] softmax statistics: context keys only (once per run) / all keys (this step)
};

struct HostIO {       // device staging for abopt_sample_host
  int N = 0, L = 0, T0 = 0;
  void* base = nullptr; size_t bytes = 0;
  float *v, *p, *res_feat, *pair_feat, *traj_v, *traj_p, *traj_prmsd, *traj_ppl;
  long long *s, *traj_s;
  uint8_t *mask_gen, *mask_res;
};

struct abopt_model {
  abopt_config cfg;
  int device = 0;
  bool finalized = false;
  std::map<std::string, KeySpec> spec;
  std::map<std::string, HostTensor> sd;
  void* wbase = nullptr; size_t wbytes = 0;
  std::vector<BlockW> blocks;
  std::vector<PairBiasParams> pb;
  std::vector<PairBiasPacked> pbp;
  // pair bias z . W_b of every layer, [slot][N][H][L queries][Lp keys].  Inside abopt_sample_* it is computed once per run for all
  // layers (z and the weights are loop invariants of the T reverse steps); elsewhere slot 0 is recomputed per block call.
  float* bias_buf = nullptr; size_t bias_slots = 0, bias_slot_floats = 0; bool bias_hoisted = false;
  bool focus_built = false;     // inside abopt_sample_*: the focus lists of mask_generate were built once for the whole run
  bool prows_built[3] = {false, false, false};      // likewise the row lists of pair_stream_kernel (mask_res / the focus list are loop invariants)
  bool ctx_built = false;       // ... and the context cache of the first block
  bool x0_built = false;        // ... and the mixer output / first-block operands of the context rows
  EpsW eps;
  DiffW diff;
  Workspace ws;
  HostIO io;
  cudaStream_t own_stream = nullptr;
  void* train_buf = nullptr; size_t train_bytes = 0;      // scratch of abopt_loss_forward
  void* design_buf = nullptr; size_t design_bytes = 0;    // scratch of abopt_design_*: context mask, v_0, p_0, res_feat, pair_feat
  void* design_io = nullptr; size_t design_io_bytes = 0;  // device staging of abopt_design_host
  // ---- training step (abopt_loss_backward): raw weights as stored by nn.Linear, gradient buffer, scratch
  struct RawHead { const float* W0; const float* b0; const float* W2; const float* b2; const float* W4; const float* b4; int nout; };
  struct RawEps { const float* Wm0; const float* bm0; const float* Wm2; const float* bm2; RawHead head[4]; const float* prm_g; const float* prm_b; } raw{};
  std::map<std::string, std::pair<size_t, size_t>> grad_off;      // trainable key -> (offset, numel) in grad_buf
  float* grad_buf = nullptr; size_t grad_floats = 0;
  void* tws = nullptr; size_t tws_bytes = 0;
  long long batch_offset = 0;                             // index of this handle's first complex in the global (unsharded) batch
};

static void add_key(abopt_model* m, const std::string& k, size_t numel, int dtype = 0, bool required = true) {
  m->spec[k] = KeySpec{numel, dtype, required};
}

static void build_spec(abopt_model* m) {
  const int T1 = m->cfg.num_steps + 1;
  const int scope = m->cfg.scope;
  if (scope != ABOPT_SCOPE_ENCODER) {
    add_key(m, "eps_net.current_sequence_embedding.weight", 25 * F);
    add_key(m, "eps_net.res_feat_mixer.0.weight", F * 2 * F); add_key(m, "eps_net.res_feat_mixer.0.bias", F);
    add_key(m, "eps_net.res_feat_mixer.2.weight", F * F);     add_key(m, "eps_net.res_feat_mixer.2.bias", F);
  }
  for (int l = 0; l < m->cfg.num_layers; ++l) {
    const std::string p = "eps_net.encoder.blocks." + std::to_string(l) + ".";
    add_key(m, p + "spatial_coef", H);
    add_key(m, p + "proj_query.weight", H * D * F); add_key(m, p + "proj_key.weight", H * D * F);
    add_key(m, p + "proj_value.weight", H * D * F); add_key(m, p + "proj_pair_bias.weight", H * C);
    add_key(m, p + "proj_query_point.weight", H * P * 3 * F); add_key(m, p + "proj_key_point.weight", H * P * 3 * F);
    add_key(m, p + "proj_value_point.weight", H * P * 3 * F);
    add_key(m, p + "out_transform.weight", F * NFEAT); add_key(m, p + "out_transform.bias", F);
    add_key(m, p + "layer_norm_1.gamma", F); add_key(m, p + "layer_norm_1.beta", F);
    add_key(m, p + "layer_norm_2.gamma", F); add_key(m, p + "layer_norm_2.beta", F);
    for (int i : {0, 2, 4}) {
      add_key(m, p + "mlp_transition." + std::to_string(i) + ".weight", F * F);
      add_key(m, p + "mlp_transition." + std::to_string(i) + ".bias", F);
    }
  }
  if (scope == ABOPT_SCOPE_ENCODER) return;
  const char* heads[3] = {"eps_crd_net", "eps_rot_net", "eps_seq_net"};
  const int nout[3] = {3, 3, NAA};
  for (int hh = 0; hh < 3; ++hh) {
    const std::string p = std::string("eps_net.") + heads[hh] + ".";
    add_key(m, p + "0.weight", F * (F + 3)); add_key(m, p + "0.bias", F);
    add_key(m, p + "2.weight", F * F);       add_key(m, p + "2.bias", F);
    add_key(m, p + "4.weight", (size_t)nout[hh] * F); add_key(m, p + "4.bias", nout[hh]);
  }
  if (m->cfg.has_prmsd) {
    const std::string p = "eps_net.prmsd_predictor.";
    add_key(m, p + "layer_norm.gamma", F + 3); add_key(m, p + "layer_norm.beta", F + 3);
    add_key(m, p + "linear_1.weight", F * (F + 3)); add_key(m, p + "linear_1.bias", F);
    add_key(m, p + "linear_2.weight", F * F);       add_key(m, p + "linear_2.bias", F);
    add_key(m, p + "linear_3.weight", (size_t)m->cfg.prmsd_bins * F); add_key(m, p + "linear_3.bias", m->cfg.prmsd_bins);
    if (scope == ABOPT_SCOPE_FULL) add_key(m, "prmsd.tobin.offset", m->cfg.prmsd_bins, 0, false);
  }
  if (scope != ABOPT_SCOPE_FULL) return;
  for (const char* mod : {"trans_rot", "trans_pos", "trans_seq"})
    for (const char* k : {"betas", "alpha_bars", "alphas", "sigmas", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod"})
      add_key(m, std::string(mod) + ".var_sched." + k, T1);
  for (const char* tab : {"fwd", "inv"}) {
    const std::string p = std::string("trans_rot.angular_distrib_") + tab + ".";
    add_key(m, p + "stddevs", T1); add_key(m, p + "approx_flag", T1, 1);
    add_key(m, p + "X", (size_t)T1 * NBINS); add_key(m, p + "Y", (size_t)T1 * NBINS);
  }
  add_key(m, "position_mean", 3); add_key(m, "position_scale", 1);
  add_key(m, "_dummy", 0, 0, false); add_key(m, "trans_rot._dummy", 0, 0, false);
}

extern "C" int abopt_version(void) { return 100; }
// Test hook for the tcgen05 3xTF32 GEMM: D[M][N] = A[M][K] * B[N][K]^T (+ bias), all device pointers.
extern "C" int abopt_debug_gemm3x(int device, int M, int N, int K, const float* A, const float* B, const float* bias, float* D, void* stream) {
  if (M < 1 || N < 4 || N % 4 || K < 32 || K % 32) return fail(ABOPT_ERR_ARG, "need N % 4 == 0 and K % 32 == 0");
  DeviceGuard g(device);
  CUDA_TRY(tc_init());
  cudaStream_t st = (cudaStream_t)stream;
  // exactly as the model uses it: the raw fp32 operands serve as the "hi" planes (the tensor core truncates to tf32)
  float *Al = nullptr, *Bl = nullptr;
  CUDA_TRY(cudaMalloc(&Al, (size_t)M * K * 4));
  {
    const cudaError_t e2 = cudaMalloc(&Bl, (size_t)N * K * 4);
    if (e2 != cudaSuccess) { cudaFree(Al); return fail(ABOPT_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e2)); }
  }
  launch_lo(A, Al, (size_t)M * K, st);
  launch_lo(B, Bl, (size_t)N * K, st);
  const bool ok = launch_gemm3x_plain(M, N, K, A, Al, K, B, Bl, K, D, N, bias, st);
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(Al); cudaFree(Bl);
  if (!ok) return fail(ABOPT_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  if (e != cudaSuccess) return fail(ABOPT_ERR_CUDA, std::string("gemm3x: ") + cudaGetErrorString(e));
  CHECK_LAUNCH();
  return ABOPT_OK;
}
// Debug: SM-clock timestamps of the phases of CTA (0,0,0) of the last attn_logits_tc_kernel launch (synchronises).
extern "C" int abopt_debug_clocks(long long* out16) {
  if (!out16) return fail(ABOPT_ERR_ARG, "null argument");
  CUDA_TRY(cudaDeviceSynchronize());
  attn_debug_clocks(out16);                 // slots 0..9 (legacy one-tile-per-CTA logits kernel only)
  tail_debug_clocks(out16 + 10);            // slots 10..14: outT_tail_kernel CTA 0: start, phase 1 done, LN1 done, MLP done, end
  return ABOPT_OK;
}
// Debug hook: copy one internal workspace tensor of the LAST block call into `dst` (device memory, room for `max_floats`):
// which = 0 QA, 1 KB, 2 rq, 3 rk, 4 VT, 5 pair bias slot 0, 6 alpha, 7 feat.  Returns the number of floats copied in *numel.
extern "C" int abopt_debug_copy(abopt_model* m, int which, float* dst, size_t max_floats, size_t* numel, void* stream) {
  if (!m || !dst || !numel) return fail(ABOPT_ERR_ARG, "null argument");
  Workspace& w = m->ws;
  if (!w.base) return fail(ABOPT_ERR_STATE, "no workspace yet");
  const size_t M = (size_t)w.N * w.L;
  const float* src = nullptr; size_t n = 0;
  switch (which) {
    case 0: src = w.op.QA; n = M * H * 64; break;
    case 1: src = w.op.KB; n = M * H * 64; break;
    case 2: src = w.op.rq; n = M * H; break;
    case 3: src = w.op.rk; n = M * H; break;
    case 4: src = w.op.VT; n = (size_t)w.N * H * 64 * w.Lp; break;
    case 5: src = m->bias_buf; n = m->bias_slot_floats; break;
    case 6: src = w.alpha; n = (size_t)w.NB * H * w.L * w.Lp; break;
    case 7: src = w.feat; n = M * NFEAT; break;
    default: return fail(ABOPT_ERR_ARG, "unknown tensor");
  }
  if (!src) return fail(ABOPT_ERR_STATE, "tensor not allocated");
  if (n > max_floats) n = max_floats;
  DeviceGuard g(m->device);
  CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  *numel = n;
  return ABOPT_OK;
}

extern "C" int abopt_profile_enable(int on) {
  for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  g_prof_on = on != 0;
  return ABOPT_OK;
}
extern "C" int abopt_profile_collect(double* ms_per_kind, uint64_t* launches_per_kind, int n_kinds) {
  if (!ms_per_kind || !launches_per_kind || n_kinds < KK_COUNT) return fail(ABOPT_ERR_ARG, "need room for all kernel kinds");
  for (int k = 0; k < n_kinds; ++k) { ms_per_kind[k] = 0.0; launches_per_kind[k] = 0; }
  for (auto& r : g_prof) {
    CUDA_TRY(cudaEventSynchronize(r.b));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, r.a, r.b));
    ms_per_kind[r.kind] += ms; launches_per_kind[r.kind] += 1;
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  g_prof.clear();
  return ABOPT_OK;
}
extern "C" const char* abopt_last_error(void) { return g_err.c_str(); }
extern "C" uint64_t abopt_kernel_launch_count(void) { return (uint64_t)g_launches; }

extern "C" int abopt_model_create(const abopt_config* cfg, int device, abopt_model** out) {
  if (!cfg || !out) return fail(ABOPT_ERR_ARG, "null argument");
  if (cfg->num_layers < 1 || cfg->num_layers > 64) return fail(ABOPT_ERR_ARG, "num_layers out of range");
  if (cfg->num_steps < 2 || cfg->num_steps > 4096) return fail(ABOPT_ERR_ARG, "num_steps out of range");
  if (cfg->scope < 0 || cfg->scope > 2) return fail(ABOPT_ERR_ARG, "bad scope");
  if (cfg->has_prmsd && (cfg->prmsd_bins < 2 || cfg->prmsd_bins > 42)) return fail(ABOPT_ERR_ARG, "prmsd_bins must be in [2, 42]");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(ABOPT_ERR_ARG, "no such CUDA device");
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(ABOPT_ERR_CUDA, std::string("libabopt_b200 needs a B200-class GPU (sm_100); found ") + prop.name);
  DeviceGuard g(device);
  CUDA_TRY(linear_kernels_init());
  CUDA_TRY(attn_kernels_init());
  CUDA_TRY(tc_init());
  CUDA_TRY(pair_stream_init());
  CUDA_TRY(attn_tc_init());
  CUDA_TRY(aggr_tc_init());
  CUDA_TRY(tail_tc_init());
  CUDA_TRY(backward_kernels_init());
  abopt_model* m = new abopt_model();
  m->cfg = *cfg;
  m->device = device;
  build_spec(m);
  {
    const cudaError_t e = cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete m; return fail(ABOPT_ERR_CUDA, std::string("cudaStreamCreateWithFlags: ") + cudaGetErrorString(e)); }
  }
  *out = m;
  return ABOPT_OK;
}

extern "C" void abopt_model_destroy(abopt_model* m) {
  if (!m) return;
  DeviceGuard g(m->device);
  if (m->wbase) cudaFree(m->wbase);
  if (m->ws.base) cudaFree(m->ws.base);
  if (m->io.base) cudaFree(m->io.base);
  if (m->bias_buf) cudaFree(m->bias_buf);
  if (m->train_buf) cudaFree(m->train_buf);
  if (m->design_buf) cudaFree(m->design_buf);
  if (m->design_io) cudaFree(m->design_io);
  if (m->grad_buf) cudaFree(m->grad_buf);
  if (m->tws) cudaFree(m->tws);
  if (m->own_stream) cudaStreamDestroy(m->own_stream);
  delete m;
}

extern "C" int abopt_model_set_batch_offset(abopt_model* m, int64_t first_complex) {
  if (!m) return fail(ABOPT_ERR_ARG, "null model");
  if (first_complex < 0) return fail(ABOPT_ERR_ARG, "first_complex must be >= 0");
  m->batch_offset = (long long)first_complex;
  return ABOPT_OK;
}

extern "C" int abopt_model_set_tensor(abopt_model* m, const char* key, const void* data, size_t numel, int dtype, int on_device) {
  if (!m || !key) return fail(ABOPT_ERR_ARG, "null argument");
  auto it = m->spec.find(key);
  if (it == m->spec.end()) return fail(ABOPT_ERR_KEY, std::string("unexpected state-dict key: ") + key);
  if (it->second.numel != numel) return fail(ABOPT_ERR_KEY, std::string("size mismatch for ") + key + ": expected " +
                                             std::to_string(it->second.numel) + " elements, got " + std::to_string(numel));
  if (it->second.dtype != dtype) return fail(ABOPT_ERR_KEY, std::string("dtype mismatch for ") + key);
  const size_t esz = dtype == 0 ? 4 : (dtype == 1 ? 1 : 8);
  HostTensor& t = m->sd[key];
  t.bytes.resize(numel * esz); t.numel = numel; t.dtype = dtype; t.set = true;
  if (numel) {
    if (!data) return fail(ABOPT_ERR_ARG, "null data");
    if (on_device) { DeviceGuard g(m->device); CUDA_TRY(cudaMemcpy(t.bytes.data(), data, numel * esz, cudaMemcpyDeviceToHost)); }
    else memcpy(t.bytes.data(), data, numel * esz);
  }
  m->finalized = false;
  return ABOPT_OK;
}

// bump allocator over a host image of the packed weights
struct Packer {
  std::vector<unsigned char> img;
  size_t put(const void* src, size_t bytes) {
    size_t off = (img.size() + 255) & ~size_t(255);
    img.resize(off + bytes);
    memcpy(img.data() + off, src, bytes);
    return off;
  }
  size_t put(const std::vector<float>& v) { return put(v.data(), v.size() * sizeof(float)); }
};

static const float* F32(abopt_model* m, const std::string& k) { return reinterpret_cast<const float*>(m->sd[k].bytes.data()); }
static std::vector<float> transpose(const float* w, int rows, int cols, int pad_rows_to = 0) {   // [rows][cols] -> [cols][max(rows,pad)]
  const int R = pad_rows_to > rows ? pad_rows_to : rows;
  std::vector<float> o((size_t)cols * R, 0.f);
  for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) o[(size_t)c * R + r] = w[(size_t)r * cols + c];
  return o;
}

static std::vector<float> lo_plane(const float* w, size_t n) {      // x - trunc_tf32(x), see k_tc.cu
  std::vector<float> o(n);
  for (size_t i = 0; i < n; ++i) {
    uint32_t u; memcpy(&u, &w[i], 4); u &= 0xFFFFE000u;
    float h; memcpy(&h, &u, 4);
    o[i] = w[i] - h;
  }
  return o;
}

extern "C" int abopt_model_finalize(abopt_model* m) {
  if (!m) return fail(ABOPT_ERR_ARG, "null model");
  for (auto& kv : m->spec)
    if (kv.second.required && !m->sd[kv.first].set) return fail(ABOPT_ERR_STATE, "missing state-dict key: " + kv.first);
  DeviceGuard g(m->device);
  Packer pk;
  struct Fix { const float** slot; size_t off; };
  std::vector<Fix> fix;
  auto reg = [&](const float** slot, size_t off) { fix.push_back({slot, off}); };
  const int NL = m->cfg.num_layers;
  m->blocks.assign(NL, BlockW{});
  m->pb.assign(NL, PairBiasParams{});
  m->pbp.assign(NL, PairBiasPacked{});
  for (int l = 0; l < NL; ++l) {
    const std::string p = "eps_net.encoder.blocks." + std::to_string(l) + ".";
    BlockW& b = m->blocks[l];
    std::vector<float> wcat((size_t)NPROJ * F);
    size_t o = 0;
    for (const char* nm : {"proj_query", "proj_key", "proj_value", "proj_query_point", "proj_key_point", "proj_value_point"}) {
      const HostTensor& t = m->sd[p + nm + ".weight"];
      memcpy(wcat.data() + o, t.bytes.data(), t.numel * 4); o += t.numel;
    }
    reg(&b.Wcat, pk.put(wcat));
    reg(&b.Wcat_lo, pk.put(lo_plane(wcat.data(), wcat.size())));
    {
      const HostTensor& wo = m->sd[p + "out_transform.weight"];                 // [128][1824]
      reg(&b.Wout, pk.put(wo.bytes.data(), wo.numel * 4));
      reg(&b.Wout_lo, pk.put(lo_plane(reinterpret_cast<const float*>(wo.bytes.data()), wo.numel)));
    }
    const float* wb = F32(m, p + "proj_pair_bias.weight");         // [12][64]
    std::vector<float> wbt = transpose(wb, H, C);                    // [64][12]
    reg(&b.Wb, pk.put(wbt));
    reg(&b.Wb_raw, pk.put(wb, H * C * 4));
    reg(&b.sc_raw, pk.put(F32(m, p + "spatial_coef"), H * 4));
    memcpy(m->pb[l].Wb, wbt.data(), sizeof(float) * C * H);
    for (int c = 0; c < C; ++c)
      for (int hp = 0; hp < H / 2; ++hp) m->pbp[l].w[hp / 3][c][hp % 3] = make_float2(wb[(2 * hp) * C + c], wb[(2 * hp + 1) * C + c]);
    std::vector<float> coef(H);
    const float* sc = F32(m, p + "spatial_coef");
    for (int h = 0; h < H; ++h) {                                    // ga.py:109-111
      const float gamma = sc[h] > 20.f ? sc[h] : log1pf(expf(sc[h]));            // F.softplus (threshold 20)
      coef[h] = (-1.f * gamma * (float)std::sqrt(2.0 / (9.0 * P))) / 2.f;
      m->pb[l].coef[h] = coef[h];
    }
    reg(&b.coef, pk.put(coef));
    reg(&b.Wout_t, pk.put(transpose(F32(m, p + "out_transform.weight"), F, NFEAT)));
    reg(&b.bout, pk.put(F32(m, p + "out_transform.bias"), F * 4));
    reg(&b.ln1_g, pk.put(F32(m, p + "layer_norm_1.gamma"), F * 4)); reg(&b.ln1_b, pk.put(F32(m, p + "layer_norm_1.beta"), F * 4));
    reg(&b.ln2_g, pk.put(F32(m, p + "layer_norm_2.gamma"), F * 4)); reg(&b.ln2_b, pk.put(F32(m, p + "layer_norm_2.beta"), F * 4));
    reg(&b.W1_t, pk.put(transpose(F32(m, p + "mlp_transition.0.weight"), F, F))); reg(&b.b1, pk.put(F32(m, p + "mlp_transition.0.bias"), F * 4));
    reg(&b.W2_t, pk.put(transpose(F32(m, p + "mlp_transition.2.weight"), F, F))); reg(&b.b2, pk.put(F32(m, p + "mlp_transition.2.bias"), F * 4));
    reg(&b.W3_t, pk.put(transpose(F32(m, p + "mlp_transition.4.weight"), F, F))); reg(&b.b3, pk.put(F32(m, p + "mlp_transition.4.bias"), F * 4));
    {
      std::vector<float> wm((size_t)3 * F * F);
      int li = 0;
      for (const char* nm : {"mlp_transition.0.weight", "mlp_transition.2.weight", "mlp_transition.4.weight"})
        memcpy(wm.data() + (size_t)(li++) * F * F, F32(m, p + nm), sizeof(float) * F * F);
      reg(&b.Wmlp, pk.put(wm));
      reg(&b.Wmlp_lo, pk.put(lo_plane(wm.data(), wm.size())));
    }
  }
  EpsW& e = m->eps;
  e = EpsW{};
  DiffW& d = m->diff;
  d = DiffW{};
  std::vector<size_t> flag_off(2, 0);
  if (m->cfg.scope != ABOPT_SCOPE_ENCODER) {
  e.has_prmsd = m->cfg.has_prmsd; e.prmsd_bins = m->cfg.prmsd_bins;
  reg(&e.emb, pk.put(F32(m, "eps_net.current_sequence_embedding.weight"), 25 * F * 4));
  reg(&e.Wm0_t, pk.put(transpose(F32(m, "eps_net.res_feat_mixer.0.weight"), F, 2 * F)));
  reg(&e.bm0, pk.put(F32(m, "eps_net.res_feat_mixer.0.bias"), F * 4));
  reg(&e.Wm2_t, pk.put(transpose(F32(m, "eps_net.res_feat_mixer.2.weight"), F, F)));
  reg(&e.bm2, pk.put(F32(m, "eps_net.res_feat_mixer.2.bias"), F * 4));
  auto pack_head = [&](HeadW& hw, const std::string& w0, const std::string& b0, const std::string& w2, const std::string& b2,
                       const std::string& w4, const std::string& b4, int nout) {
    const float* W0 = F32(m, w0);                                    // [128][131]
    std::vector<float> w0t((size_t)F * F), w0e((size_t)3 * F);
    for (int n = 0; n < F; ++n) {
      for (int k = 0; k < F; ++k) w0t[(size_t)k * F + n] = W0[(size_t)n * (F + 3) + k];
      for (int q = 0; q < 3; ++q) w0e[(size_t)q * F + n] = W0[(size_t)n * (F + 3) + F + q];
    }
    reg(&hw.W0_t, pk.put(w0t)); reg(&hw.W0_ext, pk.put(w0e));
    reg(&hw.b0, pk.put(F32(m, b0), F * 4));
    reg(&hw.W2_t, pk.put(transpose(F32(m, w2), F, F))); reg(&hw.b2, pk.put(F32(m, b2), F * 4));
    reg(&hw.W4_t, pk.put(transpose(F32(m, w4), nout, F, F)));
    std::vector<float> b4p(F, 0.f);
    memcpy(b4p.data(), F32(m, b4), nout * 4);
    reg(&hw.b4, pk.put(b4p));
  };
  pack_head(e.crd, "eps_net.eps_crd_net.0.weight", "eps_net.eps_crd_net.0.bias", "eps_net.eps_crd_net.2.weight",
            "eps_net.eps_crd_net.2.bias", "eps_net.eps_crd_net.4.weight", "eps_net.eps_crd_net.4.bias", 3);
  pack_head(e.rot, "eps_net.eps_rot_net.0.weight", "eps_net.eps_rot_net.0.bias", "eps_net.eps_rot_net.2.weight",
            "eps_net.eps_rot_net.2.bias", "eps_net.eps_rot_net.4.weight", "eps_net.eps_rot_net.4.bias", 3);
  pack_head(e.seq, "eps_net.eps_seq_net.0.weight", "eps_net.eps_seq_net.0.bias", "eps_net.eps_seq_net.2.weight",
            "eps_net.eps_seq_net.2.bias", "eps_net.eps_seq_net.4.weight", "eps_net.eps_seq_net.4.bias", NAA);
  {
    auto rawt = [&](const float** slot, const std::string& k) { const HostTensor& t = m->sd[k]; reg(slot, pk.put(t.bytes.data(), t.numel * 4)); };
    rawt(&m->raw.Wm0, "eps_net.res_feat_mixer.0.weight"); rawt(&m->raw.bm0, "eps_net.res_feat_mixer.0.bias");
    rawt(&m->raw.Wm2, "eps_net.res_feat_mixer.2.weight"); rawt(&m->raw.bm2, "eps_net.res_feat_mixer.2.bias");
    const char* hn[3] = {"eps_net.eps_crd_net.", "eps_net.eps_rot_net.", "eps_net.eps_seq_net."};
    const int no[3] = {3, 3, NAA};
    for (int hh = 0; hh < 3; ++hh) {
      auto& rh = m->raw.head[hh];
      rh.nout = no[hh];
      rawt(&rh.W0, std::string(hn[hh]) + "0.weight"); rawt(&rh.b0, std::string(hn[hh]) + "0.bias");
      rawt(&rh.W2, std::string(hn[hh]) + "2.weight"); rawt(&rh.b2, std::string(hn[hh]) + "2.bias");
      rawt(&rh.W4, std::string(hn[hh]) + "4.weight"); rawt(&rh.b4, std::string(hn[hh]) + "4.bias");
    }
    if (m->cfg.has_prmsd) {
      const std::string p = "eps_net.prmsd_predictor.";
      auto& rh = m->raw.head[3];
      rh.nout = m->cfg.prmsd_bins;
      rawt(&rh.W0, p + "linear_1.weight"); rawt(&rh.b0, p + "linear_1.bias"); rawt(&rh.W2, p + "linear_2.weight"); rawt(&rh.b2, p + "linear_2.bias");
      rawt(&rh.W4, p + "linear_3.weight"); rawt(&rh.b4, p + "linear_3.bias");
      rawt(&m->raw.prm_g, p + "layer_norm.gamma"); rawt(&m->raw.prm_b, p + "layer_norm.beta");
    }
  }
  if (m->cfg.has_prmsd) {
    const std::string p = "eps_net.prmsd_predictor.";
    pack_head(e.prm, p + "linear_1.weight", p + "linear_1.bias", p + "linear_2.weight", p + "linear_2.bias",
              p + "linear_3.weight", p + "linear_3.bias", m->cfg.prmsd_bins);
    reg(&e.prm_ln_g, pk.put(F32(m, p + "layer_norm.gamma"), (F + 3) * 4));
    reg(&e.prm_ln_b, pk.put(F32(m, p + "layer_norm.beta"), (F + 3) * 4));
  }
  }
  if (m->cfg.scope == ABOPT_SCOPE_FULL) {
  const int T1 = m->cfg.num_steps + 1;
  d.num_steps = m->cfg.num_steps; d.obj_pred_x0 = m->cfg.obj_pred_x0;
  d.prmsd_min = m->cfg.prmsd_min; d.prmsd_max = m->cfg.prmsd_max;
  memcpy(d.pos_mean, F32(m, "position_mean"), 12);
  d.pos_scale = F32(m, "position_scale")[0];
  reg(&d.betas, pk.put(F32(m, "trans_pos.var_sched.betas"), T1 * 4));
  reg(&d.alpha_bars, pk.put(F32(m, "trans_pos.var_sched.alpha_bars"), T1 * 4));
  reg(&d.alphas, pk.put(F32(m, "trans_pos.var_sched.alphas"), T1 * 4));
  reg(&d.sigmas, pk.put(F32(m, "trans_pos.var_sched.sigmas"), T1 * 4));
  reg(&d.sqrt_recip_ab, pk.put(F32(m, "trans_pos.var_sched.sqrt_recip_alphas_cumprod"), T1 * 4));
  reg(&d.sqrt_recipm1_ab, pk.put(F32(m, "trans_pos.var_sched.sqrt_recipm1_alphas_cumprod"), T1 * 4));
  reg(&d.alpha_bars_rot, pk.put(F32(m, "trans_rot.var_sched.alpha_bars"), T1 * 4));
  reg(&d.alpha_bars_seq, pk.put(F32(m, "trans_seq.var_sched.alpha_bars"), T1 * 4));
  for (int tab = 0; tab < 2; ++tab) {
    const std::string p = std::string("trans_rot.angular_distrib_") + (tab == 0 ? "fwd" : "inv") + ".";
    reg(&d.ang_X[tab], pk.put(F32(m, p + "X"), (size_t)T1 * NBINS * 4));
    const float* Y = F32(m, p + "Y");
    reg(&d.ang_Y[tab], pk.put(Y, (size_t)T1 * NBINS * 4));
    std::vector<float> cdf((size_t)T1 * NBINS, 1.f);               // inverse-CDF table for fast mode
    for (int t = 0; t < T1; ++t) {
      double tot = 0.0;
      for (int k = 0; k < NBINS - 1; ++k) tot += (double)Y[(size_t)t * NBINS + k];
      double run = 0.0;
      for (int k = 0; k < NBINS - 1; ++k) {
        run += (double)Y[(size_t)t * NBINS + k];
        cdf[(size_t)t * NBINS + k] = tot > 0.0 ? (float)(run / tot) : (float)(k + 1) / (float)(NBINS - 1);
      }
      cdf[(size_t)t * NBINS + NBINS - 2] = 2.f;                      // sentinel: search never runs past the last bin
    }
    reg(&d.ang_cdf[tab], pk.put(cdf));
    reg(&d.ang_std[tab], pk.put(F32(m, p + "stddevs"), T1 * 4));
    flag_off[tab] = pk.put(m->sd[p + "approx_flag"].bytes.data(), T1);
  }
  }
  if (m->wbase) { CUDA_TRY(cudaFree(m->wbase)); m->wbase = nullptr; }
  m->wbytes = pk.img.size();
  CUDA_TRY(cudaMalloc(&m->wbase, m->wbytes));
  CUDA_TRY(cudaMemcpy(m->wbase, pk.img.data(), m->wbytes, cudaMemcpyHostToDevice));
  for (auto& f : fix) *f.slot = reinterpret_cast<const float*>(static_cast<unsigned char*>(m->wbase) + f.off);
  if (m->cfg.scope == ABOPT_SCOPE_FULL)
    for (int tab = 0; tab < 2; ++tab) d.ang_flag[tab] = static_cast<const uint8_t*>(m->wbase) + flag_off[tab];
  // gradient buffer: one slot per trainable tensor; the six projection weights of a block are contiguous in the order of
  // Wcat (q | k | v | query points | key points | value points) so that one (2016 x 128) weight-gradient GEMM fills them
  m->grad_off.clear();
  size_t goff = 0;
  auto gslot = [&](const std::string& k) { const size_t n = m->sd[k].numel; m->grad_off[k] = {goff, n}; goff += (n + 3) & ~size_t(3); };
  for (int l = 0; l < NL; ++l) {
    const std::string p = "eps_net.encoder.blocks." + std::to_string(l) + ".";
    for (const char* nm : {"proj_query", "proj_key", "proj_value", "proj_query_point", "proj_key_point", "proj_value_point"}) gslot(p + nm + ".weight");
  }
  for (auto& kv : m->spec) {
    const std::string& k = kv.first;
    if (!kv.second.required || kv.second.dtype != 0 || m->grad_off.count(k)) continue;
    if (k.rfind("trans_", 0) == 0 || k.rfind("position_", 0) == 0 || k.rfind("prmsd.", 0) == 0 || k.rfind("_dummy", 0) == 0) continue;
    gslot(k);
  }
  if (m->grad_buf) { CUDA_TRY(cudaFree(m->grad_buf)); m->grad_buf = nullptr; }
  m->grad_floats = goff;
  if (goff) { CUDA_TRY(cudaMalloc(&m->grad_buf, goff * sizeof(float))); CUDA_TRY(cudaMemset(m->grad_buf, 0, goff * sizeof(float))); }
  m->finalized = true;
  return ABOPT_OK;
}

// ------------------------------------------------------------------------------------------ workspace
static int chunk_size(int N, int L, int Lp) {
  const char* env = getenv("ABOPT_CHUNK");
  if (env && atoi(env) > 0) return atoi(env) < N ? atoi(env) : N;
  // Complexes per pass of the attention kernels.  Measured on B200 (profiles/): one pass over the whole batch beats
  // L2-sized chunks -- the per-launch ramp-up / tail of the persistent kernels costs more than the L2 misses on alpha.
  // The only cap is memory: the attention weights of a chunk (H * L * Lp floats per complex) stay below 2 GiB.
  const double per = (double)H * L * Lp * 4.0;
  int nb = (int)(2.0 * 1024 * 1024 * 1024 / per);
  if (nb < 1) nb = 1;
  if (nb >= N) return N;
  const int nch = (N + nb - 1) / nb;
  return (N + nch - 1) / nch;
}

static int ensure_workspace(abopt_model* m, int N, int L) {
  Workspace& w = m->ws;
  if (w.base && w.N >= N && w.L == L) return ABOPT_OK;
  if (w.base) { CUDA_TRY(cudaFree(w.base)); w.base = nullptr; }
  const int Lp = (L + 7) & ~7;      // row pitch of the attention tensors: 32-byte aligned rows (256-bit stores)
  const int NB = chunk_size(N, L, Lp);
  const size_t M = (size_t)N * L;
  const int bins = m->cfg.has_prmsd ? m->cfg.prmsd_bins : 1;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~size_t(255); return o; };
  const size_t oR = take(M * 9 * 4), oP = take(M * 3 * 4), oXa = take(M * F * 4), oXb = take(M * F * 4),
               oFeat = take(M * NFEAT * 4),
               oAl = take((size_t)NB * H * L * Lp * 4), oPr = take(M * bins * 4), oPl = take((size_t)N * bins * 4),
               oMp = take(M * 4), oVn = take(M * 3 * 4), oEp = take(M * 3 * 4), oCd = take(M * NAA * 4), oRn = take(M * 9 * 4),
               oBi = take(M * 4), oTv = take((size_t)N * 8), oXal = take(M * F * 4), oXbl = take(M * F * 4), oXil = take(M * F * 4),
               oQa = take(M * H * 64 * 4), oQl = take(256),
               oKb = take(M * H * 64 * 4), oKl = take(256), oRq = take(M * H * 4), oRk = take(M * H * 4),
               oVt = take((size_t)N * H * 64 * Lp * 4),
               oFc = take(M * 4), oFr = take(M * 4), oFw = take((size_t)N * (L / 64 + 2) * 8), oFn = take(64), oFs = take((size_t)N * 8),
               oFx = take(M * F * 4), oFm = take(M), oPr0 = take(M * 16), oPr1 = take(M * 16), oPr2 = take(M * 16), oPrc = take(64),
               oCc = take(M * H * C * 4), oCs = take(M * H * 8), oCt = take(M * H * 8),
               oX0 = take(M * F * 4), oX0l = take(M * F * 4), oX0c = take(M * F * 4), oX0cl = take(M * F * 4),
               oQa0 = take(M * H * 64 * 4), oKb0 = take(M * H * 64 * 4), oRq0 = take(M * H * 4), oRk0 = take(M * H * 4),
               oVt0 = take((size_t)N * H * 64 * Lp * 4);
  CUDA_TRY(cudaMalloc(&w.base, off));
  CUDA_TRY(cudaMemset(w.base, 0, off));        // the padding rows / columns of the packed attention operands must stay zero
  unsigned char* b = static_cast<unsigned char*>(w.base);
  w.Rbuf = (float*)(b + oR); w.pnorm = (float*)(b + oP); w.xa = (float*)(b + oXa); w.xb = (float*)(b + oXb);
  w.feat = (float*)(b + oFeat); w.alpha = (float*)(b + oAl);
  w.prmsd_rows = (float*)(b + oPr); w.prmsd_logits = (float*)(b + oPl); w.maxprob = (float*)(b + oMp);
  w.v_net = (float*)(b + oVn); w.eps_pos = (float*)(b + oEp); w.c_den = (float*)(b + oCd); w.R_next = (float*)(b + oRn);
  w.bin_idx = (int*)(b + oBi); w.tvec_scratch = (long long*)(b + oTv);
  w.xa_lo = (float*)(b + oXal); w.xb_lo = (float*)(b + oXbl); w.xin_lo = (float*)(b + oXil);
  w.op = AttnOperands{(float*)(b + oQa), (float*)(b + oQl), (float*)(b + oKb), (float*)(b + oKl), (float*)(b + oRq), (float*)(b + oRk),
                      (float*)(b + oVt)};
  w.focus = Focus{(int*)(b + oFc), (int*)(b + oFr), (int2*)(b + oFw), (int*)(b + oFn), (int*)(b + oFs), (float*)(b + oFx), (uint8_t*)(b + oFm)};
  w.prows[0] = PairRows{(int4*)(b + oPr0), (int*)(b + oPrc)};
  w.prows[1] = PairRows{(int4*)(b + oPr1), (int*)(b + oPrc) + 2};
  w.prows[2] = PairRows{(int4*)(b + oPr2), (int*)(b + oPrc) + 4};
  w.ctx_cache = (float*)(b + oCc); w.ctx_stats = (float2*)(b + oCs); w.step_stats = (float2*)(b + oCt);
  w.x0 = (float*)(b + oX0); w.x0_lo = (float*)(b + oX0l); w.x0_c = (float*)(b + oX0c); w.x0_c_lo = (float*)(b + oX0cl);
  w.op0 = AttnOperands{(float*)(b + oQa0), (float*)(b + oQl), (float*)(b + oKb0), (float*)(b + oKl), (float*)(b + oRq0), (float*)(b + oRk0),
                       (float*)(b + oVt0)};
  w.N = N; w.L = L; w.Lp = Lp; w.NB = NB; w.bytes = off;
  return ABOPT_OK;
}

extern "C" size_t abopt_workspace_bytes(const abopt_model* m) { return m ? m->ws.bytes : 0; }

static int check_ready(abopt_model* m, int N, int L, int need_scope = ABOPT_SCOPE_ENCODER) {
  if (!m) return fail(ABOPT_ERR_ARG, "null model");
  // scope order of capability: FULL (0) > EPSNET (2) > ENCODER (1)
  const int have = m->cfg.scope;
  const bool ok = (need_scope == ABOPT_SCOPE_ENCODER) || (need_scope == ABOPT_SCOPE_EPSNET && have != ABOPT_SCOPE_ENCODER) ||
                  (need_scope == ABOPT_SCOPE_FULL && have == ABOPT_SCOPE_FULL);
  if (!ok) return fail(ABOPT_ERR_STATE, "model scope does not include this operation");
  if (!m->finalized) return fail(ABOPT_ERR_STATE, "model not finalised (call abopt_model_finalize after loading every tensor)");
  if (N < 1 || L < 1) return fail(ABOPT_ERR_ARG, "N and L must be positive");
  if (L > ABOPT_MAX_L) return fail(ABOPT_ERR_ARG, "L exceeds ABOPT_MAX_L (" + std::to_string(ABOPT_MAX_L) + ")");
  if ((size_t)N * L > (size_t)1 << 30) return fail(ABOPT_ERR_ARG, "N*L too large");
  // the Philox counters are keyed by the 32-bit row index in the global batch (k_step.cu)
  if ((unsigned long long)(m->batch_offset + N) * (unsigned long long)L > 0xFFFFFFFFull)
    return fail(ABOPT_ERR_ARG, "batch offset + N exceeds the 32-bit global row index");
  return ABOPT_OK;
}

// ------------------------------------------------------------------------------------------ encoder
// room for `slots` layers of pair bias
static int ensure_pair_inputs(abopt_model* m, int N, int L, const float* z, size_t slots) {
  (void)z;
  const size_t slot_floats = (size_t)N * H * L * ((L + 7) & ~7);
  if (!m->bias_buf || m->bias_slot_floats != slot_floats || m->bias_slots < slots) {
    if (m->bias_buf) { CUDA_TRY(cudaFree(m->bias_buf)); m->bias_buf = nullptr; }
    CUDA_TRY(cudaMalloc(&m->bias_buf, slots * slot_floats * sizeof(float)));
    m->bias_slots = slots; m->bias_slot_floats = slot_floats;
  }
  return ABOPT_OK;
}

// one GABlock: x_in -> x_out (may not alias); feat/alpha taps optional.
// x_lo = tf32 "lo" plane of x (nullptr: computed here); x_lo_out = where to put the lo plane of x_out (may be nullptr)
// fc != nullptr ("focus", last block inside the sampling loop): only the rows listed in fc are produced, x_out is COMPACT
static int run_block(abopt_model* m, int layer, int N, int L, const float* R, const float* t, const float* x, const float* x_lo,
                     const float* z, const uint8_t* mask, float* x_out, float* x_lo_out, float* alpha_tap, cudaStream_t st,
                     const Focus* fc = nullptr, const uint8_t* ctx_cache_gen = nullptr) {
  Workspace& w = m->ws;
  const int M = N * L;
  const BlockW& bw = m->blocks[layer];
  int rc = ensure_pair_inputs(m, N, L, z, m->bias_hoisted ? (size_t)m->cfg.num_layers : 1); if (rc) return rc;
  const float* bias = m->bias_buf + (m->bias_hoisted ? (size_t)layer * m->bias_slot_floats : 0);
  if (!m->bias_hoisted && !launch_pair_bias(N, 0, N, L, w.Lp, z, m->pbp[layer], m->bias_buf, st))
    return fail(ABOPT_ERR_ARG, "pair_bias_kernel: L too large for shared memory");
  // the six input projections: tcgen05 3xTF32 GEMM (x raw = "hi" plane, x_lo = "lo" plane)
  if (x_lo == nullptr) { launch_lo(x, w.xin_lo, (size_t)M * F, st); x_lo = w.xin_lo; }
  for (int b0 = 0; b0 < N; b0 += w.NB) {
    const int nb = (N - b0 < w.NB) ? (N - b0) : w.NB;
    // one pass = as many complexes as the alpha buffer holds (normally the whole batch, see chunk_size): projections ->
    // packed attention operands -> alpha -> aggregates; every tensor between two kernels travels through HBM / L2 once
    // Loop invariance of the first block inside the sampling loop (context cache, k_pair.cu): everything a context residue feeds
    // into it is the same at every step.  Its packed operands live in their own buffers (w.op0: the other layers reuse w.op), so
    // after the first step only the generated rows are projected; the context part of the pair aggregate is computed once per
    // run and only the generated keys' share of z is streamed per step.
    const bool ctx = layer == 0 && ctx_cache_gen != nullptr && nb == N && fc == nullptr;
    const AttnOperands& ops = ctx ? w.op0 : w.op;
    if (ctx && m->ctx_built) {
      // x = w.x0 here; w.x0_c / w.x0_c_lo hold its generated rows, compact (mixer_kernel)
      if (!launch_proj_pack(N * L, L, w.Lp, w.x0_c, w.x0_c_lo, bw.Wcat, bw.Wcat_lo, R, t, bw.coef, ops, st, w.focus.rows, w.focus.count))
        return fail(ABOPT_ERR_CUDA, "cuTensorMapEncodeTiled failed (proj)");
    } else {
      const size_t r0 = (size_t)b0 * L, o64 = (size_t)b0 * H * L * 64, o1 = (size_t)b0 * H * L, ov = (size_t)b0 * H * 64 * w.Lp;
      // (QA_lo / KB_lo are placeholders: every logits kernel builds the lo planes of its operands on chip)
      const AttnOperands opc{ops.QA + o64, ops.QA_lo, ops.KB + o64, ops.KB_lo, ops.rq + o1, ops.rk + o1, ops.VT + ov};
      if (!launch_proj_pack(nb * L, L, w.Lp, x + r0 * F, x_lo + r0 * F, bw.Wcat, bw.Wcat_lo, R + r0 * 9, t + r0 * 3, bw.coef, opc, st))
        return fail(ABOPT_ERR_CUDA, "cuTensorMapEncodeTiled failed (proj)");
    }
    if (ctx && !m->ctx_built) {
      // once per run: softmax over the context keys alone (+ its statistics) and the pair aggregate of every query row with it
      if (!launch_attn_logits_tc(nb, b0, N, L, w.Lp, ops, bias, mask, w.alpha, st, nullptr, nullptr, ctx_cache_gen, w.ctx_stats))
        return fail(ABOPT_ERR_CUDA, "attn_logits_tc launch failed");
      if (!m->prows_built[0]) { launch_pair_rows_build(nb, b0, L, mask, nullptr, w.prows[0], st); m->prows_built[0] = true; }
      if (!launch_pair_stream(nb, b0, L, w.Lp, z, w.alpha, w.ctx_cache, w.prows[0], st, H * C))
        return fail(ABOPT_ERR_CUDA, "cuTensorMapEncodeTiled failed (pair)");
      launch_pair_rows_build(nb, b0, L, mask, w.focus.cidx, w.prows[2], st, /*compact=*/false);
      m->ctx_built = true;
    }
    // logits (node + spatial + pair bias, scaled, masked) and softmax on the tensor cores -> alpha
    if (!launch_attn_logits_tc(nb, b0, N, L, w.Lp, ops, bias, mask, w.alpha, st, fc ? fc->windows : nullptr, fc ? fc->count : nullptr,
                               nullptr, ctx ? w.step_stats : nullptr))
      return fail(ABOPT_ERR_CUDA, "attn_logits_tc launch failed");
    if (ctx) {
      // generated query rows: all of their z rows; context query rows: cached context part + the generated keys
      if (!launch_pair_stream(nb, b0, L, w.Lp, z, w.alpha, w.feat, w.prows[2], st, NFEAT, /*partial=*/true))
        return fail(ABOPT_ERR_CUDA, "cuTensorMapEncodeTiled failed (pair)");
      launch_ctx_delta(N, L, w.Lp, z, mask, w.alpha, w.ctx_cache, w.ctx_stats, w.step_stats, w.focus.cidx, w.focus.rows, w.focus.scratch,
                       w.focus.count, w.feat, st);
    } else {
      const int which = fc ? 1 : 0;
      if (!m->prows_built[which]) launch_pair_rows_build(nb, b0, L, mask, fc ? fc->cidx : nullptr, w.prows[which], st);
      if (m->bias_hoisted && nb == N) m->prows_built[which] = true;      // the sampling loop: the masks are loop invariants
      if (!launch_pair_stream(nb, b0, L, w.Lp, z, w.alpha, w.feat, w.prows[which], st, NFEAT, /*partial=*/fc != nullptr))
        return fail(ABOPT_ERR_CUDA, "cuTensorMapEncodeTiled failed (pair)");
    }
    if (!launch_aggr_tc(nb, b0, N, L, w.Lp, w.alpha, ops.VT, R, t, w.feat, st, fc ? fc->windows : nullptr,
                        fc ? fc->count : nullptr, fc ? fc->cidx : nullptr))
      return fail(ABOPT_ERR_CUDA, "aggr_tc launch failed");
    if (alpha_tap) launch_alpha_tap(nb, b0, L, w.Lp, w.alpha, alpha_tap, st);
  }
  if (x_out) {
    // out_transform (K = 1824) on the tensor cores, then mask / residual / LN / MLP / LN, one kernel
    if (fc) {
      launch_focus_gather(fc->rows, fc->count, x, mask, fc->x_c, fc->mask_c, st);
      if (!launch_outT_tail(M, w.feat, fc->x_c, fc->mask_c, bw, x_out, nullptr, st, fc->count))
        return fail(ABOPT_ERR_CUDA, "cuTensorMapEncodeTiled failed (tail)");
    } else if (!launch_outT_tail(M, w.feat, x, mask, bw, x_out, x_lo_out, st)) {
      return fail(ABOPT_ERR_CUDA, "cuTensorMapEncodeTiled failed (tail)");
    }
  }
  CHECK_LAUNCH();
  return ABOPT_OK;
}

extern "C" int abopt_ga_block_forward(abopt_model* m, int layer, int N, int L, const float* R, const float* t,
                                      const float* x, const float* z, const uint8_t* mask, float* x_out, void* stream) {
  int rc = check_ready(m, N, L); if (rc) return rc;
  if (layer < 0 || layer >= m->cfg.num_layers) return fail(ABOPT_ERR_ARG, "layer out of range");
  if (!R || !t || !x || !z || !mask || !x_out) return fail(ABOPT_ERR_ARG, "null tensor");
  DeviceGuard g(m->device);
  rc = ensure_workspace(m, N, L); if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  float* dst = (x_out == x) ? m->ws.xa : x_out;
  rc = run_block(m, layer, N, L, R, t, x, nullptr, z, mask, dst, nullptr, nullptr, st); if (rc) return rc;
  if (dst != x_out) CUDA_TRY(cudaMemcpyAsync(x_out, dst, (size_t)N * L * F * 4, cudaMemcpyDeviceToDevice, st));
  return ABOPT_OK;
}

extern "C" int abopt_ga_block_taps(abopt_model* m, int layer, int N, int L, const float* R, const float* t,
                                   const float* x, const float* z, const uint8_t* mask, float* alpha, float* feat, void* stream) {
  int rc = check_ready(m, N, L); if (rc) return rc;
  if (layer < 0 || layer >= m->cfg.num_layers) return fail(ABOPT_ERR_ARG, "layer out of range");
  if (!R || !t || !x || !z || !mask) return fail(ABOPT_ERR_ARG, "null tensor");
  DeviceGuard g(m->device);
  rc = ensure_workspace(m, N, L); if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  rc = run_block(m, layer, N, L, R, t, x, nullptr, z, mask, nullptr, nullptr, alpha, st); if (rc) return rc;
  if (feat) CUDA_TRY(cudaMemcpyAsync(feat, m->ws.feat, (size_t)N * L * NFEAT * 4, cudaMemcpyDeviceToDevice, st));
  return ABOPT_OK;
}

// all layers; result lands in *result (one of the two workspace ping-pong buffers)
static int run_encoder(abopt_model* m, int N, int L, const float* R, const float* t, const float* x, const float* x_lo,
                       const float* z, const uint8_t* mask, float** result, cudaStream_t st, const Focus* fc = nullptr,
                       const uint8_t* ctx_cache_gen = nullptr) {
  Workspace& w = m->ws;
  const float* cur = x;
  const float* cur_lo = x_lo;
  float* bufs[2] = {w.xa, w.xb};
  float* bufs_lo[2] = {w.xa_lo, w.xb_lo};
  int which = (x == w.xa) ? 1 : 0;
  for (int l = 0; l < m->cfg.num_layers; ++l) {
    const bool last = l == m->cfg.num_layers - 1;
    int rc = run_block(m, l, N, L, R, t, cur, cur_lo, z, mask, bufs[which], bufs_lo[which], nullptr, st, last ? fc : nullptr,
                       l == 0 ? ctx_cache_gen : nullptr);
    if (rc) return rc;
    cur = bufs[which]; cur_lo = bufs_lo[which]; which ^= 1;
  }
  *result = const_cast<float*>(cur);
  return ABOPT_OK;
}

extern "C" int abopt_ga_encoder_forward(abopt_model* m, int N, int L, const float* R, const float* t, const float* x,
                                        const float* z, const uint8_t* mask, float* x_out, void* stream) {
  int rc = check_ready(m, N, L); if (rc) return rc;
  if (!R || !t || !x || !z || !mask || !x_out) return fail(ABOPT_ERR_ARG, "null tensor");
  DeviceGuard g(m->device);
  rc = ensure_workspace(m, N, L); if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  float* res = nullptr;
  rc = run_encoder(m, N, L, R, t, x, nullptr, z, mask, &res, st); if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(x_out, res, (size_t)N * L * F * 4, cudaMemcpyDeviceToDevice, st));
  return ABOPT_OK;
}

// ------------------------------------------------------------------------------------------ EpsilonNet
// p_ang != null: positions arrive in Angstrom and are normalised on the fly into ws.pnorm
static int run_eps_net(abopt_model* m, int N, int L, const float* v_t, const float* p_t, const float* p_ang,
                       const long long* s_t, const float* res_feat, const float* pair_feat, const float* beta, int beta_stride,
                       const uint8_t* mask_gen, const uint8_t* mask_res, float* v_next, float* R_next, float* eps_pos,
                       float* c_den, float* prmsd_logits, cudaStream_t st, bool focus = false) {
  Workspace& w = m->ws;
  const int M = N * L;
  // "focus": the caller consumes the outputs on generated residues only (sampling loop, no pRMSD head), so the last
  // GABlock and the heads run on those rows alone.  Results on the consumed rows are unchanged.
  // ABOPT_NO_FOCUS=1 computes every row (read per call: the parity tests run both ways in one process)
  const char* nf = getenv("ABOPT_NO_FOCUS");
  const bool focus_off = nf && nf[0] == '1';
  const Focus* fc = nullptr;      // last block + heads on the generated rows (models without the pRMSD head)
  const Focus* hf = nullptr;      // crd / rot / seq heads on the generated rows (every model); the pRMSD head needs all rows
  if (focus && !focus_off && m->cfg.num_layers >= 1) {
    if (!m->focus_built)
      launch_focus_build(N, L, mask_gen, w.focus.cidx, w.focus.rows, w.focus.windows, w.focus.count, w.focus.scratch, st);
    if (m->bias_hoisted) m->focus_built = true;           // the sampling loop: mask_generate is a loop invariant
    hf = &w.focus;
    if (!m->cfg.has_prmsd && w.NB >= N) fc = &w.focus;
  }
  // context invariance inside the sampling loop (everything a context residue feeds into the mixer and the first block is a loop
  // invariant there): with the focus lists built.  ABOPT_NO_CTXCACHE=1 disables it (read per call, like ABOPT_NO_FOCUS).
  const char* nc = getenv("ABOPT_NO_CTXCACHE");
  const bool ctx_on = hf != nullptr && m->bias_hoisted && m->focus_built && w.NB >= N && m->cfg.num_layers >= 2 && p_ang != nullptr &&
                      !(nc && nc[0] == '1');
  float* x0 = ctx_on ? w.x0 : w.xa;
  float* x0_lo = ctx_on ? w.x0_lo : w.xa_lo;
  if (ctx_on && m->x0_built) {
    // generated rows only: in place, and compact for the first block's projection
    launch_mixer(M, res_feat, s_t, v_t, m->eps, x0, w.Rbuf, p_ang, w.pnorm, m->diff.pos_mean, m->diff.pos_scale, x0_lo, st,
                 w.focus.rows, w.focus.count, w.x0_c, w.x0_c_lo);
  } else {
    launch_mixer(M, res_feat, s_t, v_t, m->eps, x0, w.Rbuf, p_ang, w.pnorm, m->diff.pos_mean, m->diff.pos_scale, x0_lo, st);
    if (ctx_on) m->x0_built = true;
  }
  const float* tpos = p_ang ? w.pnorm : p_t;
  float* enc = nullptr;
  int rc = run_encoder(m, N, L, w.Rbuf, tpos, x0, x0_lo, pair_feat, mask_res, &enc, st, fc, ctx_on ? mask_gen : nullptr);
  if (rc) return rc;
  launch_heads(M, L, enc, beta, beta_stride, w.Rbuf, v_t, mask_gen, m->eps, v_next, R_next, eps_pos, c_den, w.prmsd_rows,
               m->cfg.has_prmsd ? (prmsd_logits ? prmsd_logits : w.prmsd_logits) : nullptr, st, hf ? hf->rows : nullptr,
               hf ? hf->count : nullptr, /*x_compact=*/fc != nullptr);
  CHECK_LAUNCH();
  return ABOPT_OK;
}

extern "C" int abopt_eps_net_forward(abopt_model* m, int N, int L, const float* v_t, const float* p_t, const int64_t* s_t,
                                     const float* res_feat, const float* pair_feat, const float* beta,
                                     const uint8_t* mask_generate, const uint8_t* mask_res, float* v_next, float* R_next,
                                     float* eps_pos, float* c_denoised, float* prmsd_logits, void* stream) {
  int rc = check_ready(m, N, L, ABOPT_SCOPE_EPSNET); if (rc) return rc;
  if (!v_t || !p_t || !s_t || !res_feat || !pair_feat || !beta || !mask_generate || !mask_res || !v_next || !eps_pos || !c_denoised)
    return fail(ABOPT_ERR_ARG, "null tensor");
  DeviceGuard g(m->device);
  rc = ensure_workspace(m, N, L); if (rc) return rc;
  return run_eps_net(m, N, L, v_t, p_t, nullptr, (const long long*)s_t, res_feat, pair_feat, beta, 1, mask_generate, mask_res,
                     v_next, R_next, eps_pos, c_denoised, prmsd_logits, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------ transitions
extern "C" int abopt_rot_denoise(abopt_model* m, int N, int L, const float* v_t, const float* v_net, const uint8_t* mask_generate,
                                 const int64_t* t, const float* u, const float* expo_ang, const float* unif_ang,
                                 const float* gauss_ang, float* v_out, void* stream) {
  int rc = check_ready(m, N, L, ABOPT_SCOPE_FULL); if (rc) return rc;
  if (!v_t || !v_net || !mask_generate || !t || !u || !expo_ang || !unif_ang || !gauss_ang || !v_out) return fail(ABOPT_ERR_ARG, "null tensor");
  DeviceGuard g(m->device);
  rc = ensure_workspace(m, N, L); if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  launch_angle_argmax(N * L, L, (const long long*)t, 0, m->diff.ang_Y[1], expo_ang, mask_generate, m->ws.bin_idx, st);
  NoisePtrs nz{u, unif_ang, gauss_ang, nullptr, nullptr, m->ws.bin_idx};
  launch_rot_denoise(N * L, L, v_t, v_net, mask_generate, (const long long*)t, nz, m->diff, v_out, st);
  CHECK_LAUNCH();
  return ABOPT_OK;
}

extern "C" int abopt_pos_pred_noise_from_start(abopt_model* m, int N, int L, const float* p_t, const float* p_0,
                                               const uint8_t* mask_generate, const int64_t* t, float* eps_out, void* stream) {
  int rc = check_ready(m, N, L, ABOPT_SCOPE_FULL); if (rc) return rc;
  if (!p_t || !p_0 || !mask_generate || !t || !eps_out) return fail(ABOPT_ERR_ARG, "null tensor");
  DeviceGuard g(m->device);
  launch_pos(N * L, L, 0, p_t, p_0, mask_generate, (const long long*)t, nullptr, m->diff, eps_out, (cudaStream_t)stream);
  CHECK_LAUNCH();
  return ABOPT_OK;
}

extern "C" int abopt_pos_denoise(abopt_model* m, int N, int L, const float* p_t, const float* eps_p, const uint8_t* mask_generate,
                                 const int64_t* t, const float* z_pos, float* p_out, void* stream) {
  int rc = check_ready(m, N, L, ABOPT_SCOPE_FULL); if (rc) return rc;
  if (!p_t || !eps_p || !mask_generate || !t || !z_pos || !p_out) return fail(ABOPT_ERR_ARG, "null tensor");
  DeviceGuard g(m->device);
  launch_pos(N * L, L, 1, p_t, eps_p, mask_generate, (const long long*)t, z_pos, m->diff, p_out, (cudaStream_t)stream);
  CHECK_LAUNCH();
  return ABOPT_OK;
}

extern "C" int abopt_seq_denoise(abopt_model* m, int N, int L, const int64_t* s_t, const float* c0_pred, const uint8_t* mask_generate,
                                 const int64_t* t, const float* expo_seq, float* post, int64_t* s_out, void* stream) {
  int rc = check_ready(m, N, L, ABOPT_SCOPE_FULL); if (rc) return rc;
  if (!s_t || !c0_pred || !mask_generate || !t || !expo_seq || !post || !s_out) return fail(ABOPT_ERR_ARG, "null tensor");
  DeviceGuard g(m->device);
  launch_seq_denoise(N * L, L, (const long long*)s_t, c0_pred, mask_generate, (const long long*)t, expo_seq, m->diff, post,
                     (long long*)s_out, (cudaStream_t)stream);
  CHECK_LAUNCH();
  return ABOPT_OK;
}

// ------------------------------------------------------------------------------------------ sampling loop
static int run_init(abopt_model* m, int N, int L, const float* v, const float* p, const long long* s, const uint8_t* mask_generate,
                    uint32_t flags, int opt_step, uint64_t seed, const abopt_init_noise* init_noise, float* v_out, float* p_out,
                    long long* s_out, float* prmsd_out, float* ppl_out, cudaStream_t st) {
  Workspace& w = m->ws;
  const int M = N * L;
  const bool optimize = opt_step > 0;
  const int T0 = optimize ? opt_step : m->cfg.num_steps;
  InitArgs ia{};
  ia.M = M; ia.L = L; ia.T0 = T0;
  ia.sample_structure = (flags & ABOPT_SAMPLE_STRUCTURE) ? 1 : 0; ia.sample_sequence = (flags & ABOPT_SAMPLE_SEQUENCE) ? 1 : 0;
  ia.optimize = optimize ? 1 : 0;
  ia.has_prmsd = (m->cfg.has_prmsd && prmsd_out && ppl_out) ? 1 : 0;
  ia.v = v; ia.p_ang = p; ia.s = s; ia.mask_gen = mask_generate;
  ia.v_out = v_out; ia.p_out_ang = p_out; ia.s_out = s_out; ia.prmsd_out = prmsd_out; ia.ppl_out = ppl_out;
  ia.seed = seed; ia.row0 = (uint32_t)(m->batch_offset * L);
  if (init_noise) {
    if (!optimize) {
      if (!init_noise->g4 || !init_noise->gp || !init_noise->s_rand) return fail(ABOPT_ERR_ARG, "init_noise for sample() needs g4, gp, s_rand");
      ia.g4 = init_noise->g4; ia.gp = init_noise->gp; ia.s_rand = (const long long*)init_noise->s_rand;
    } else {
      const abopt_step_noise* a = init_noise->add;
      if (!a || !a->u || !a->expo_ang || !a->unif_ang || !a->gauss_ang || !a->z_pos || !a->expo_seq)
        return fail(ABOPT_ERR_ARG, "init_noise for optimize() needs a full abopt_step_noise");
      launch_angle_argmax(M, L, nullptr, T0, m->diff.ang_Y[0], a->expo_ang, mask_generate, w.bin_idx, st);
      ia.add = NoisePtrs{a->u, a->unif_ang, a->gauss_ang, a->z_pos, a->expo_seq, w.bin_idx};
    }
  }
  launch_init(ia, m->diff, st);
  CHECK_LAUNCH();
  return ABOPT_OK;
}

// one iteration of the loop body: traj[t] -> traj[t-1]
static int run_step(abopt_model* m, int N, int L, int t, bool optimize, uint32_t flags, uint64_t seed, const float* v_t,
                    const float* p_t_ang, const long long* s_t, const float* res_feat, const float* pair_feat,
                    const uint8_t* mask_generate, const uint8_t* mask_res, const abopt_step_noise* nz, float* v_out, float* p_out,
                    long long* s_out, float* prmsd_out, float* ppl_out, cudaStream_t st) {
  Workspace& w = m->ws;
  const int M = N * L;
  const bool prm = m->cfg.has_prmsd != 0 && prmsd_out && ppl_out;
  int rc = run_eps_net(m, N, L, v_t, nullptr, p_t_ang, s_t, res_feat, pair_feat, m->diff.betas + t, 0, mask_generate, mask_res,
                       w.v_net, w.R_next, w.eps_pos, w.c_den, nullptr, st, /*focus=*/true);
  if (rc) return rc;
  StepArgs sa{};
  sa.M = M; sa.L = L; sa.t = t;
  sa.sample_structure = (flags & ABOPT_SAMPLE_STRUCTURE) ? 1 : 0; sa.sample_sequence = (flags & ABOPT_SAMPLE_SEQUENCE) ? 1 : 0;
  sa.pred_x0 = (!optimize && m->cfg.obj_pred_x0) ? 1 : 0;         // optimize() ignores obj (dpm_full.py:351-356)
  sa.masked_ppl = optimize ? 0 : 1;                               // optimize() passes no mask (dpm_full.py:357)
  sa.v_t = v_t; sa.p_t_ang = p_t_ang; sa.s_t = s_t;
  sa.v_net = w.v_net; sa.p_pred = w.eps_pos; sa.c_den = w.c_den; sa.mask_gen = mask_generate;
  sa.v_out = v_out; sa.p_out_ang = p_out; sa.s_out = s_out;
  sa.maxprob_rows = prm ? w.maxprob : nullptr;
  sa.seed = seed; sa.row0 = (uint32_t)(m->batch_offset * L);
  if (nz) {
    if (!nz->u || !nz->expo_ang || !nz->unif_ang || !nz->gauss_ang || !nz->z_pos || !nz->expo_seq)
      return fail(ABOPT_ERR_ARG, "incomplete abopt_step_noise record");
    if (sa.sample_structure) launch_angle_argmax(M, L, nullptr, t, m->diff.ang_Y[1], nz->expo_ang, mask_generate, w.bin_idx, st);
    sa.nz = NoisePtrs{nz->u, nz->unif_ang, nz->gauss_ang, nz->z_pos, nz->expo_seq, w.bin_idx};
  }
  launch_step(sa, m->diff, st);
  if (prm)
    launch_complex_reduce(N, L, m->cfg.prmsd_bins, m->cfg.prmsd_min, m->cfg.prmsd_max, sa.masked_ppl, w.prmsd_logits, w.maxprob,
                          mask_generate, prmsd_out, ppl_out, st);
  CHECK_LAUNCH();
  return ABOPT_OK;
}

extern "C" int abopt_sample_init(abopt_model* m, int N, int L, const float* v, const float* p, const int64_t* s,
                                 const uint8_t* mask_generate, uint32_t flags, int opt_step, uint64_t seed,
                                 const abopt_init_noise* init_noise, float* v_out, float* p_out, int64_t* s_out, void* stream) {
  int rc = check_ready(m, N, L, ABOPT_SCOPE_FULL); if (rc) return rc;
  if (!v || !p || !s || !mask_generate || !v_out || !p_out || !s_out) return fail(ABOPT_ERR_ARG, "null tensor");
  if (opt_step < 0 || opt_step > m->cfg.num_steps) return fail(ABOPT_ERR_ARG, "opt_step out of range");
  DeviceGuard g(m->device);
  rc = ensure_workspace(m, N, L); if (rc) return rc;
  return run_init(m, N, L, v, p, (const long long*)s, mask_generate, flags, opt_step, seed, init_noise, v_out, p_out,
                  (long long*)s_out, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int abopt_reverse_step(abopt_model* m, int N, int L, int t, int optimize, const float* v_t, const float* p_t,
                                  const int64_t* s_t, const float* res_feat, const float* pair_feat, const uint8_t* mask_generate,
                                  const uint8_t* mask_res, uint32_t flags, uint64_t seed, const abopt_step_noise* noise,
                                  float* v_out, float* p_out, int64_t* s_out, float* prmsd_out, float* ppl_out, void* stream) {
  int rc = check_ready(m, N, L, ABOPT_SCOPE_FULL); if (rc) return rc;
  if (!v_t || !p_t || !s_t || !res_feat || !pair_feat || !mask_generate || !mask_res || !v_out || !p_out || !s_out)
    return fail(ABOPT_ERR_ARG, "null tensor");
  if (t < 1 || t > m->cfg.num_steps) return fail(ABOPT_ERR_ARG, "t out of range");
  DeviceGuard g(m->device);
  rc = ensure_workspace(m, N, L); if (rc) return rc;
  return run_step(m, N, L, t, optimize != 0, flags, seed, v_t, p_t, (const long long*)s_t, res_feat, pair_feat, mask_generate,
                  mask_res, noise, v_out, p_out, (long long*)s_out, prmsd_out, ppl_out, (cudaStream_t)stream);
}

extern "C" int abopt_sample_device(abopt_model* m, int N, int L, const float* v, const float* p, const int64_t* s,
                                   const float* res_feat, const float* pair_feat, const uint8_t* mask_generate,
                                   const uint8_t* mask_res, uint32_t flags, int opt_step, uint64_t seed,
                                   const abopt_init_noise* init_noise, const abopt_step_noise* noise, float* traj_v,
                                   float* traj_p, int64_t* traj_s, float* traj_prmsd, float* traj_ppl, void* stream) {
  int rc = check_ready(m, N, L, ABOPT_SCOPE_FULL); if (rc) return rc;
  if (!v || !p || !s || !res_feat || !pair_feat || !mask_generate || !mask_res || !traj_v || !traj_p || !traj_s)
    return fail(ABOPT_ERR_ARG, "null tensor");
  if (m->cfg.has_prmsd && (!traj_prmsd || !traj_ppl)) return fail(ABOPT_ERR_ARG, "traj_prmsd / traj_ppl required for pRMSD models");
  if (opt_step < 0 || opt_step > m->cfg.num_steps) return fail(ABOPT_ERR_ARG, "opt_step out of range");
  if ((noise == nullptr) != (init_noise == nullptr)) return fail(ABOPT_ERR_ARG, "noise and init_noise must be given together");
  DeviceGuard g(m->device);
  rc = ensure_workspace(m, N, L); if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t M = (size_t)N * L;
  const bool optimize = opt_step > 0;
  const int T0 = optimize ? opt_step : m->cfg.num_steps;
  const bool keep = (flags & ABOPT_KEEP_TRAJECTORY) != 0;
  const bool prm = m->cfg.has_prmsd != 0;
  // Slot addressing: with the full trajectory every step has its own slot; otherwise intermediate steps
  // ping-pong between slots 1 and 2 and only slots T0 and 0 are meaningful to the caller.
  if (!keep && T0 < 3) return fail(ABOPT_ERR_ARG, "opt_step < 3 requires ABOPT_KEEP_TRAJECTORY");
  auto slot_of = [&](int t) -> int { return (keep || t == T0 || t == 0) ? t : 1 + (t & 1); };
  auto V = [&](int t) { return traj_v + (size_t)slot_of(t) * M * 3; };
  auto Pp = [&](int t) { return traj_p + (size_t)slot_of(t) * M * 3; };
  auto S = [&](int t) { return (long long*)traj_s + (size_t)slot_of(t) * M; };
  auto PR = [&](int t) { return prm ? traj_prmsd + (size_t)slot_of(t) * N : nullptr; };
  auto PL = [&](int t) { return prm ? traj_ppl + (size_t)slot_of(t) * N : nullptr; };

  rc = run_init(m, N, L, v, p, (const long long*)s, mask_generate, flags, opt_step, seed, init_noise, V(T0), Pp(T0), S(T0),
                PR(T0), PL(T0), st);
  if (rc) return rc;
  // loop-invariant hoist: the pair bias z . W_b of all layers, once per run instead of once per step
  rc = ensure_pair_inputs(m, N, L, pair_feat, (size_t)m->cfg.num_layers); if (rc) return rc;
  for (int l = 0; l < m->cfg.num_layers; ++l)
    if (!launch_pair_bias(N, 0, N, L, m->ws.Lp, pair_feat, m->pbp[l], m->bias_buf + (size_t)l * m->bias_slot_floats, st))
      return fail(ABOPT_ERR_ARG, "pair_bias_kernel: L too large for shared memory");
  m->bias_hoisted = true;
  m->focus_built = false;
  m->prows_built[0] = m->prows_built[1] = false;
  m->ctx_built = false;
  m->x0_built = false;
  for (int t = T0; t >= 1 && rc == ABOPT_OK; --t)
    rc = run_step(m, N, L, t, optimize, flags, seed, V(t), Pp(t), S(t), res_feat, pair_feat, mask_generate, mask_res,
                  noise ? &noise[T0 - t] : nullptr, V(t - 1), Pp(t - 1), S(t - 1), PR(t - 1), PL(t - 1), st);
  m->bias_hoisted = false;
  m->focus_built = false;
  m->prows_built[0] = m->prows_built[1] = false;
  m->ctx_built = false;
  m->x0_built = false;
  return rc;
}

// ------------------------------------------------------------------------------------------ training forward
// FullDPM.forward without autograd: the three add_noise calls at per-complex steps t, one EpsilonNet evaluation, the loss
// dict (dpm_full.py:156-234; AbDesign :138-190).  losses_out (device, 5 floats) = rot, pos, seq, prmsd, dist.
extern "C" int abopt_loss_forward(abopt_model* m, int N, int L, const float* v_0, const float* p_0, const int64_t* s_0,
                                  const float* res_feat, const float* pair_feat, const uint8_t* mask_generate,
                                  const uint8_t* mask_res, uint32_t flags, const int64_t* t, uint64_t seed,
                                  const abopt_step_noise* noise, float* losses_out, void* stream) {
  int rc = check_ready(m, N, L, ABOPT_SCOPE_FULL); if (rc) return rc;
  if (!v_0 || !p_0 || !s_0 || !res_feat || !pair_feat || !mask_generate || !mask_res || !t || !losses_out)
    return fail(ABOPT_ERR_ARG, "null tensor");
  if (noise && (!noise->u || !noise->expo_ang || !noise->unif_ang || !noise->gauss_ang || !noise->z_pos || !noise->expo_seq))
    return fail(ABOPT_ERR_ARG, "incomplete abopt_step_noise record");
  DeviceGuard g(m->device);
  cudaStream_t st = (cudaStream_t)stream;
  rc = ensure_workspace(m, N, L); if (rc) return rc;
  Workspace& w = m->ws;
  const size_t M = (size_t)N * L;
  const size_t need = M * (3 + 3 + 3 + 6) * sizeof(float) + M * sizeof(long long) + (size_t)N * sizeof(float) + 256;
  if (m->train_bytes < need) {
    if (m->train_buf) { CUDA_TRY(cudaFree(m->train_buf)); m->train_buf = nullptr; m->train_bytes = 0; }
    CUDA_TRY(cudaMalloc(&m->train_buf, need));
    m->train_bytes = need;
  }
  long long* s_noisy = reinterpret_cast<long long*>(m->train_buf);
  float* v_noisy = reinterpret_cast<float*>(s_noisy + M);
  float* p_noisy = v_noisy + M * 3;            // Angstrom, like every position that crosses a kernel boundary
  float* zbuf = p_noisy + M * 3;
  float* rows = zbuf + M * 3;
  float* beta = rows + M * 6;
  const bool ds = (flags & ABOPT_SAMPLE_STRUCTURE) != 0, dq = (flags & ABOPT_SAMPLE_SEQUENCE) != 0;

  InitArgs ia{};
  ia.M = (int)M; ia.L = L; ia.T0 = 0;
  ia.sample_structure = ds ? 1 : 0; ia.sample_sequence = dq ? 1 : 0; ia.optimize = 1; ia.has_prmsd = 0;
  ia.v = v_0; ia.p_ang = p_0; ia.s = (const long long*)s_0; ia.mask_gen = mask_generate;
  ia.v_out = v_noisy; ia.p_out_ang = p_noisy; ia.s_out = s_noisy;
  ia.seed = seed; ia.row0 = (uint32_t)(m->batch_offset * L); ia.tvec = (const long long*)t; ia.seq_all_rows = 1; ia.z_out = ds ? zbuf : nullptr;
  ia.grad_clamp = (flags & ABOPT_GRAD_SEMANTICS) ? 1 : 0;
  if (noise) {
    if (ds) launch_angle_argmax((int)M, L, (const long long*)t, 0, m->diff.ang_Y[0], noise->expo_ang, mask_generate, w.bin_idx, st);
    ia.add = NoisePtrs{noise->u, noise->unif_ang, noise->gauss_ang, noise->z_pos, noise->expo_seq, w.bin_idx};
  }
  launch_init(ia, m->diff, st);
  launch_gather_beta(N, (const long long*)t, m->diff.betas, beta, st);
  rc = run_eps_net(m, N, L, v_noisy, nullptr, p_noisy, s_noisy, res_feat, pair_feat, beta, 1, mask_generate, mask_res,
                   w.v_net, w.R_next, w.eps_pos, w.c_den, nullptr, st);
  if (rc) return rc;
  LossArgs la{};
  la.N = N; la.L = L; la.abdock = m->cfg.has_prmsd ? 1 : 0; la.pred_x0 = m->cfg.obj_pred_x0 ? 1 : 0;
  la.has_prmsd = m->cfg.has_prmsd; la.bins = m->cfg.prmsd_bins; la.dmin = m->cfg.prmsd_min; la.dmax = m->cfg.prmsd_max;
  la.v_0 = v_0; la.p_0_ang = p_0; la.s_0 = (const long long*)s_0;
  la.p_noisy_ang = p_noisy; la.s_noisy = s_noisy; la.z = ds ? zbuf : nullptr;
  la.R_pred = w.R_next; la.p_pred = w.eps_pos; la.c_den = w.c_den; la.prmsd_logits = w.prmsd_logits;
  la.mask_gen = mask_generate; la.mask_res = mask_res; la.tvec = (const long long*)t;
  la.rows = rows; la.out = losses_out;
  launch_loss(la, m->diff, st);
  CHECK_LAUNCH();
  return ABOPT_OK;
}

static int ensure_hostio(abopt_model* m, int N, int L, int T0) {
  HostIO& io = m->io;
  if (io.base && io.N == N && io.L == L && io.T0 == T0) return ABOPT_OK;
  if (io.base) { CUDA_TRY(cudaFree(io.base)); io.base = nullptr; }
  const size_t M = (size_t)N * L, S1 = (size_t)T0 + 1;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~size_t(255); return o; };
  const size_t ov = take(M * 12), op = take(M * 12), os = take(M * 8), orf = take(M * F * 4), opf = take(M * L * C * 4),
               omg = take(M), omr = take(M), otv = take(S1 * M * 12), otp = take(S1 * M * 12), ots = take(S1 * M * 8),
               opr = take(S1 * N * 4), opl = take(S1 * N * 4);
  CUDA_TRY(cudaMalloc(&io.base, off));
  unsigned char* b = static_cast<unsigned char*>(io.base);
  io.v = (float*)(b + ov); io.p = (float*)(b + op); io.s = (long long*)(b + os); io.res_feat = (float*)(b + orf);
  io.pair_feat = (float*)(b + opf); io.mask_gen = b + omg; io.mask_res = b + omr; io.traj_v = (float*)(b + otv);
  io.traj_p = (float*)(b + otp); io.traj_s = (long long*)(b + ots); io.traj_prmsd = (float*)(b + opr); io.traj_ppl = (float*)(b + opl);
  io.N = N; io.L = L; io.T0 = T0; io.bytes = off;
  return ABOPT_OK;
}

extern "C" int abopt_sample_host(abopt_model* m, int N, int L, const float* v, const float* p, const int64_t* s,
                                 const float* res_feat, const float* pair_feat, const uint8_t* mask_generate,
                                 const uint8_t* mask_res, uint32_t flags, int opt_step, uint64_t seed, float* traj_v,
                                 float* traj_p, int64_t* traj_s, float* traj_prmsd, float* traj_ppl) {
  int rc = check_ready(m, N, L, ABOPT_SCOPE_FULL); if (rc) return rc;
  if (!v || !p || !s || !res_feat || !pair_feat || !mask_generate || !mask_res || !traj_v || !traj_p || !traj_s)
    return fail(ABOPT_ERR_ARG, "null tensor");
  if (opt_step < 0 || opt_step > m->cfg.num_steps) return fail(ABOPT_ERR_ARG, "opt_step out of range");
  DeviceGuard g(m->device);
  const int T0 = opt_step > 0 ? opt_step : m->cfg.num_steps;
  rc = ensure_hostio(m, N, L, T0); if (rc) return rc;
  HostIO& io = m->io;
  cudaStream_t st = m->own_stream;
  const size_t M = (size_t)N * L;
  CUDA_TRY(cudaMemcpyAsync(io.v, v, M * 12, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(io.p, p, M * 12, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(io.s, s, M * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(io.res_feat, res_feat, M * F * 4, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(io.mask_gen, mask_generate, M, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(io.mask_res, mask_res, M, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(io.pair_feat, pair_feat, M * L * C * 4, cudaMemcpyHostToDevice, st));
  rc = abopt_sample_device(m, N, L, io.v, io.p, (const int64_t*)io.s, io.res_feat, io.pair_feat, io.mask_gen, io.mask_res, flags,
                           opt_step, seed, nullptr, nullptr, io.traj_v, io.traj_p, (int64_t*)io.traj_s, io.traj_prmsd, io.traj_ppl, st);
  if (rc) return rc;
  const bool keep = (flags & ABOPT_KEEP_TRAJECTORY) != 0;
  const bool prm = m->cfg.has_prmsd != 0 && traj_prmsd && traj_ppl;
  auto copy_slots = [&](int first, int count) -> int {
    CUDA_TRY(cudaMemcpyAsync(traj_v + (size_t)first * M * 3, io.traj_v + (size_t)first * M * 3, (size_t)count * M * 12, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(traj_p + (size_t)first * M * 3, io.traj_p + (size_t)first * M * 3, (size_t)count * M * 12, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(traj_s + (size_t)first * M, io.traj_s + (size_t)first * M, (size_t)count * M * 8, cudaMemcpyDeviceToHost, st));
    if (prm) {
      CUDA_TRY(cudaMemcpyAsync(traj_prmsd + (size_t)first * N, io.traj_prmsd + (size_t)first * N, (size_t)count * N * 4, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaMemcpyAsync(traj_ppl + (size_t)first * N, io.traj_ppl + (size_t)first * N, (size_t)count * N * 4, cudaMemcpyDeviceToHost, st));
    }
    return ABOPT_OK;
  };
  if (keep) { rc = copy_slots(0, T0 + 1); if (rc) return rc; }
  else { rc = copy_slots(0, 1); if (rc) return rc; rc = copy_slots(T0, 1); if (rc) return rc; }
  CUDA_TRY(cudaStreamSynchronize(st));
  return ABOPT_OK;
}

// ------------------------------------------------------------------------------------------ encode + sample from atoms
// DiffusionAntibodyDesign.sample / .optimize (models/diffab.py:115-141,143-171): encode() = ResidueEmbedding + PairEmbedding +
// backbone frames, then FullDPM.sample / optimize -- one device-resident call, pair_feat (1.07 GB at B=64, L=256) never leaves
// the device.
static int grow(void** buf, size_t* have, size_t need) {
  if (*have >= need) return ABOPT_OK;
  if (*buf) { CUDA_TRY(cudaFree(*buf)); *buf = nullptr; *have = 0; }
  CUDA_TRY(cudaMalloc(buf, need));
  *have = need;
  return ABOPT_OK;
}

extern "C" int abopt_design_device(abopt_model* m, abopt_pair_embed* pe, abopt_res_embed* re, int N, int L, int num_atoms_in,
                                   const int64_t* aa, const int64_t* res_nb, const int64_t* chain_nb, const float* pos_heavyatom,
                                   const uint8_t* mask_heavyatom, const int64_t* fragment_type, const uint8_t* generate_flag,
                                   const uint8_t* mask, uint32_t flags, int opt_step, uint64_t seed, float* traj_v, float* traj_p,
                                   int64_t* traj_s, float* traj_prmsd, float* traj_ppl, void* stream) {
  int rc = check_ready(m, N, L, ABOPT_SCOPE_FULL); if (rc) return rc;
  if (!pe || !re) return fail(ABOPT_ERR_ARG, "null embedding handle");
  if (!aa || !res_nb || !chain_nb || !pos_heavyatom || !mask_heavyatom || !fragment_type || !generate_flag || !mask)
    return fail(ABOPT_ERR_ARG, "null tensor");
  if (num_atoms_in < 4) return fail(ABOPT_ERR_ARG, "need at least the backbone atoms N, CA, C, O");
  DeviceGuard g(m->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t M = (size_t)N * L;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~size_t(255); return o; };
  const size_t oc = take(M), ov = take(M * 12), op = take(M * 12), orf = take(M * F * 4), opf = take(M * L * C * 4);
  rc = grow(&m->design_buf, &m->design_bytes, off); if (rc) return rc;
  unsigned char* b = static_cast<unsigned char*>(m->design_buf);
  uint8_t* ctx = b + oc;
  float *v0 = (float*)(b + ov), *p0 = (float*)(b + op), *res_feat = (float*)(b + orf), *pair_feat = (float*)(b + opf);
  launch_design_prep((int)M, num_atoms_in, pos_heavyatom, mask_heavyatom, generate_flag, ctx, v0, p0, st);
  CHECK_LAUNCH();
  // remove_structure / remove_sequence follow sample_structure / sample_sequence (models/diffab.py:133-137)
  const uint8_t* smask = (flags & ABOPT_SAMPLE_STRUCTURE) ? ctx : nullptr;
  const uint8_t* qmask = (flags & ABOPT_SAMPLE_SEQUENCE) ? ctx : nullptr;
  rc = abopt_res_embed_forward(re, N, L, num_atoms_in, aa, res_nb, chain_nb, pos_heavyatom, mask_heavyatom, fragment_type, smask, qmask,
                               res_feat, stream);
  if (rc) return rc;
  rc = abopt_pair_embed_forward(pe, N, L, num_atoms_in, aa, res_nb, chain_nb, pos_heavyatom, mask_heavyatom, smask, qmask, pair_feat, stream);
  if (rc) return rc;
  return abopt_sample_device(m, N, L, v0, p0, aa, res_feat, pair_feat, generate_flag, mask, flags, opt_step, seed, nullptr, nullptr,
                             traj_v, traj_p, traj_s, traj_prmsd, traj_ppl, stream);
}

extern "C" int abopt_design_host(abopt_model* m, abopt_pair_embed* pe, abopt_res_embed* re, int N, int L, int num_atoms_in,
                                 const int64_t* aa, const int64_t* res_nb, const int64_t* chain_nb, const float* pos_heavyatom,
                                 const uint8_t* mask_heavyatom, const int64_t* fragment_type, const uint8_t* generate_flag,
                                 const uint8_t* mask, uint32_t flags, int opt_step, uint64_t seed, float* traj_v, float* traj_p,
                                 int64_t* traj_s, float* traj_prmsd, float* traj_ppl) {
  int rc = check_ready(m, N, L, ABOPT_SCOPE_FULL); if (rc) return rc;
  if (!aa || !res_nb || !chain_nb || !pos_heavyatom || !mask_heavyatom || !fragment_type || !generate_flag || !mask || !traj_v ||
      !traj_p || !traj_s)
    return fail(ABOPT_ERR_ARG, "null tensor");
  if (opt_step < 0 || opt_step > m->cfg.num_steps) return fail(ABOPT_ERR_ARG, "opt_step out of range");
  DeviceGuard g(m->device);
  const int T0 = opt_step > 0 ? opt_step : m->cfg.num_steps;
  const size_t M = (size_t)N * L, S1 = (size_t)T0 + 1, A = (size_t)num_atoms_in;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~size_t(255); return o; };
  const size_t oaa = take(M * 8), orn = take(M * 8), ocn = take(M * 8), opos = take(M * A * 12), oma = take(M * A), oft = take(M * 8),
               ogen = take(M), omask = take(M), otv = take(S1 * M * 12), otp = take(S1 * M * 12), ots = take(S1 * M * 8),
               opr = take(S1 * N * 4), opl = take(S1 * N * 4);
  rc = grow(&m->design_io, &m->design_io_bytes, off); if (rc) return rc;
  unsigned char* b = static_cast<unsigned char*>(m->design_io);
  cudaStream_t st = m->own_stream;
  CUDA_TRY(cudaMemcpyAsync(b + oaa, aa, M * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(b + orn, res_nb, M * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(b + ocn, chain_nb, M * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(b + opos, pos_heavyatom, M * A * 12, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(b + oma, mask_heavyatom, M * A, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(b + oft, fragment_type, M * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(b + ogen, generate_flag, M, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(b + omask, mask, M, cudaMemcpyHostToDevice, st));
  float *dtv = (float*)(b + otv), *dtp = (float*)(b + otp), *dpr = (float*)(b + opr), *dpl = (float*)(b + opl);
  int64_t* dts = (int64_t*)(b + ots);
  rc = abopt_design_device(m, pe, re, N, L, num_atoms_in, (const int64_t*)(b + oaa), (const int64_t*)(b + orn), (const int64_t*)(b + ocn),
                           (const float*)(b + opos), b + oma, (const int64_t*)(b + oft), b + ogen, b + omask, flags, opt_step, seed,
                           dtv, dtp, dts, dpr, dpl, st);
  if (rc) return rc;
  const bool keep = (flags & ABOPT_KEEP_TRAJECTORY) != 0;
  const bool prm = m->cfg.has_prmsd != 0 && traj_prmsd && traj_ppl;
  auto copy_slots = [&](int first, int count) -> int {
    CUDA_TRY(cudaMemcpyAsync(traj_v + (size_t)first * M * 3, dtv + (size_t)first * M * 3, (size_t)count * M * 12, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(traj_p + (size_t)first * M * 3, dtp + (size_t)first * M * 3, (size_t)count * M * 12, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(traj_s + (size_t)first * M, dts + (size_t)first * M, (size_t)count * M * 8, cudaMemcpyDeviceToHost, st));
    if (prm) {
      CUDA_TRY(cudaMemcpyAsync(traj_prmsd + (size_t)first * N, dpr + (size_t)first * N, (size_t)count * N * 4, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaMemcpyAsync(traj_ppl + (size_t)first * N, dpl + (size_t)first * N, (size_t)count * N * 4, cudaMemcpyDeviceToHost, st));
    }
    return ABOPT_OK;
  };
  if (keep) { rc = copy_slots(0, T0 + 1); if (rc) return rc; }
  else { rc = copy_slots(0, 1); if (rc) return rc; rc = copy_slots(T0, 1); if (rc) return rc; }
  CUDA_TRY(cudaStreamSynchronize(st));
  return ABOPT_OK;
}

// ------------------------------------------------------------------------------------------ training step: backward
// loss.backward() of AbDock/train.py:104-113 through FullDPM.forward / EpsilonNet / GAEncoder, recompute-based (csrc/k_backward.cu,
// formulas of oracle/ipa_backward.py and oracle/epsnet_backward.py).  Parameter gradients accumulate nowhere: every call overwrites
// the handle's gradient buffer (abopt_model_get_grad reads it); d res_feat / d pair_feat go to caller buffers.
namespace {
struct TrainWS {
  float *xs, *Pm, *PG, *G, *gfeat, *s1, *h, *a0, *a1, *s2, *t1, *t2, *t3, *gagg, *glog, *part, *scr, *gx, *hcat, *ghcat, *cat0, *gcat, *GO, *GP,
      *lnb, *orot, *stats, *glogit, *tmp;
  size_t scr_floats;
};
}
static int ensure_train_ws(abopt_model* m, int N, int L, int Lp, TrainWS& t) {
  const size_t M = (size_t)N * L;
  const int bins = m->cfg.has_prmsd ? m->cfg.prmsd_bins : 1;
  size_t off = 0;
  auto take = [&](size_t floats) { size_t o = off; off = (off + floats * 4 + 255) & ~size_t(255); return o; };
  // scratch: the split weight-gradient partials (64 x NPROJ x F) or, in block_backward, the lo plane of an (M, F) activation
  // followed by a transposed weight pair (hi | lo) of at most NPROJ x F each -- whichever is larger
  size_t scr_floats = (size_t)64 * NPROJ * F;
  if (scr_floats < M * F + (size_t)2 * NPROJ * F) scr_floats = M * F + (size_t)2 * NPROJ * F;
  const size_t o_xs = take(M * F * (m->cfg.num_layers + 1)), o_P = take(M * NPROJ), o_PG = take(M * 864), o_G = take(M * NPROJ),
               o_gf = take(M * NFEAT), o_s1 = take(M * F), o_h = take(M * F), o_a0 = take(M * F), o_a1 = take(M * F), o_s2 = take(M * F),
               o_t1 = take(M * F), o_t2 = take(M * F), o_t3 = take(M * F), o_ga = take(M * 288), o_gl = take((size_t)N * H * L * Lp),
               o_pt = take(M * 780), o_sc = take(scr_floats), o_gx = take(M * F), o_hc = take(M * (F + 3)), o_gh = take(M * (F + 3)),
               o_c0 = take(M * 2 * F), o_gc = take(M * 2 * F), o_go = take(M * 26), o_gp = take(M * bins), o_ln = take(M * (F + 3)),
               o_or = take(M * 4), o_st = take(16), o_gg = take((size_t)N * bins), o_tm = take(4096);
  if (m->tws_bytes < off) {
    if (m->tws) { CUDA_TRY(cudaFree(m->tws)); m->tws = nullptr; m->tws_bytes = 0; }
    CUDA_TRY(cudaMalloc(&m->tws, off));
    m->tws_bytes = off;
  }
  unsigned char* b = static_cast<unsigned char*>(m->tws);
  auto P_ = [&](size_t o) { return reinterpret_cast<float*>(b + o); };
  t = TrainWS{P_(o_xs), P_(o_P), P_(o_PG), P_(o_G), P_(o_gf), P_(o_s1), P_(o_h), P_(o_a0), P_(o_a1), P_(o_s2), P_(o_t1), P_(o_t2), P_(o_t3),
              P_(o_ga), P_(o_gl), P_(o_pt), P_(o_sc), P_(o_gx), P_(o_hc), P_(o_gh), P_(o_c0), P_(o_gc), P_(o_go), P_(o_gp), P_(o_ln), P_(o_or),
              P_(o_st), P_(o_gg), P_(o_tm), scr_floats};
  return ABOPT_OK;
}
static float* grad_of(abopt_model* m, const std::string& key) {
  auto it = m->grad_off.find(key);
  return it == m->grad_off.end() ? nullptr : m->grad_buf + it->second.first;
}
__global__ void coef_grad_kernel(const float* __restrict__ gc, const float* __restrict__ sc, float* __restrict__ out) {
  const int h = threadIdx.x;
  if (h >= H) return;
  const float cP = sqrtf(2.f / (9.f * P)) / 2.f;                       // ga.py:109-111: coef = -softplus(sc) cP
  out[h] = gc[h] * (-cP) * (1.f / (1.f + expf(-sc[h])));               // d softplus = sigmoid
}
__global__ void strided_copy_kernel(int M, int K, const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)M * K) return;
  const int r = (int)(i / K), c = (int)(i - (size_t)r * K);
  dst[(size_t)r * ldd + c] = src[(size_t)r * lds + c];
}

// one GABlock, backward.  x = block input; gx: in = d loss / d block output, out = d loss / d x.  d z written (or accumulated) to dz.
static int block_backward(abopt_model* m, TrainWS& T, int l, int N, int L, const float* R, const float* t, const float* x, const float* z,
                          const uint8_t* mask, float* gx, float* dz, bool dz_accumulate, cudaStream_t st) {
  Workspace& w = m->ws;
  const int M = N * L;
  const BlockW& bw = m->blocks[l];
  const std::string p = "eps_net.encoder.blocks." + std::to_string(l) + ".";
  if (w.NB < N) return fail(ABOPT_ERR_ARG, "backward: the batch does not fit one attention pass (alpha > 2 GiB)");
  // ---- recompute: alpha and the aggregate by the forward kernels, the plain projections, the tail's activations
  int rc = run_block(m, l, N, L, R, t, x, nullptr, z, mask, nullptr, nullptr, nullptr, st); if (rc) return rc;
  // (the two big recompute GEMMs on the tensor cores, 3xTF32 like the forward: x's lo plane is w.xin_lo since run_block; the lo
  // plane of the aggregate goes through T.gfeat, which is not written before the tail backward below)
  if (!launch_gemm3x_plain(M, NPROJ, F, x, w.xin_lo, F, bw.Wcat, bw.Wcat_lo, F, T.Pm, NPROJ, nullptr, st))
    return fail(ABOPT_ERR_CUDA, "cuTensorMapEncodeTiled failed (backward: projections)");
  bwd_points_global(M, T.Pm, R, t, T.PG, st);
  const float *W1 = bw.Wmlp, *W2 = bw.Wmlp + F * F, *W3 = bw.Wmlp + 2 * F * F;
  launch_lo(w.feat, T.gfeat, (size_t)M * NFEAT, st);
  if (!launch_gemm3x_plain(M, F, NFEAT, w.feat, T.gfeat, NFEAT, bw.Wout, bw.Wout_lo, NFEAT, T.t1, F, bw.bout, st))      // y = out_transform(feat)
    return fail(ABOPT_ERR_CUDA, "cuTensorMapEncodeTiled failed (backward: out_transform)");
  bwd_add_ln_fwd(M, x, T.t1, mask, bw.ln1_g, bw.ln1_b, T.s1, T.h, st);                                  // s1 = x + mask y ; h = LN1
  bwd_gemm_nt(M, F, F, T.h, F, false, W1, F, bw.b1, T.a0, F, st);
  bwd_gemm_nt(M, F, F, T.a0, F, true, W2, F, bw.b2, T.a1, F, st);
  bwd_gemm_nt(M, F, F, T.a1, F, true, W3, F, bw.b3, T.t1, F, st);                                       // m3
  bwd_add_ln_fwd(M, T.h, T.t1, nullptr, nullptr, nullptr, T.s2, nullptr, st);                           // s2 = h + m3
  // ---- tail backward (ga.py:173-178)
  bwd_ln(M, gx, nullptr, T.s2, bw.ln2_g, T.t2, T.t1, nullptr, st);                                      // t2 = g_s2, t1 = g * xhat
  bwd_colsum(M, F, T.t1, F, nullptr, 0, grad_of(m, p + "layer_norm_2.gamma"), T.scr, st, 1.f);
  bwd_colsum(M, F, gx, F, nullptr, 0, grad_of(m, p + "layer_norm_2.beta"), T.scr, st, 1.f);
  bwd_wgrad(M, F, F, T.t2, F, T.a1, F, true, grad_of(m, p + "mlp_transition.4.weight"), T.scr, T.scr_floats, st);
  bwd_colsum(M, F, T.t2, F, nullptr, 0, grad_of(m, p + "mlp_transition.4.bias"), T.scr, st, 1.f);
  bwd_gemm_nn(M, F, F, T.t2, F, W3, F, T.t3, F, false, st);
  bwd_relu((size_t)M * F, T.t3, T.a1, st);                                                              // t3 = g_r1
  bwd_wgrad(M, F, F, T.t3, F, T.a0, F, true, grad_of(m, p + "mlp_transition.2.weight"), T.scr, T.scr_floats, st);
  bwd_colsum(M, F, T.t3, F, nullptr, 0, grad_of(m, p + "mlp_transition.2.bias"), T.scr, st, 1.f);
  bwd_gemm_nn(M, F, F, T.t3, F, W2, F, T.t1, F, false, st);
  bwd_relu((size_t)M * F, T.t1, T.a0, st);                                                              // t1 = g_r0
  bwd_wgrad(M, F, F, T.t1, F, T.h, F, false, grad_of(m, p + "mlp_transition.0.weight"), T.scr, T.scr_floats, st);
  bwd_colsum(M, F, T.t1, F, nullptr, 0, grad_of(m, p + "mlp_transition.0.bias"), T.scr, st, 1.f);
  bwd_gemm_nn(M, F, F, T.t1, F, W1, F, T.t3, F, false, st);                                             // t3 = g_h (through the MLP)
  bwd_ln(M, T.t3, T.t2, T.s1, bw.ln1_g, gx, T.t1, T.a1, st);                                            // gx = g_s1 ; a1 = g_h + g_s2
  bwd_colsum(M, F, T.t1, F, nullptr, 0, grad_of(m, p + "layer_norm_1.gamma"), T.scr, st, 1.f);
  bwd_colsum(M, F, T.a1, F, nullptr, 0, grad_of(m, p + "layer_norm_1.beta"), T.scr, st, 1.f);
  bwd_add_mask(M, F, gx, nullptr, mask, T.t2, st);                                                      // t2 = g_s1 * mask
  {
    // d feat = g_y W_out on the tensor cores: A = g_y (lo plane in the scratch block), B = W_out^T [1824][128] hi | lo
    float* a_lo = T.scr; float* wt_h = a_lo + (size_t)M * F; float* wt_l = wt_h + (size_t)F * NFEAT;
    launch_lo(T.t2, a_lo, (size_t)M * F, st);
    launch_transpose_split(bw.Wout, F, NFEAT, wt_h, wt_l, st);
    if (!launch_gemm3x_plain(M, NFEAT, F, T.t2, a_lo, F, wt_h, wt_l, F, T.gfeat, NFEAT, nullptr, st))
      return fail(ABOPT_ERR_CUDA, "cuTensorMapEncodeTiled failed (backward: d feat)");
  }
  bwd_wgrad(M, F, NFEAT, T.t2, F, w.feat, NFEAT, false, grad_of(m, p + "out_transform.weight"), T.scr, T.scr_floats, st);
  bwd_colsum(M, F, T.t2, F, nullptr, 0, grad_of(m, p + "out_transform.bias"), T.scr, st, 1.f);
  // ---- aggregate, softmax, logits (ga.py:81-147)
  bwd_aggregate(M, T.gfeat, w.feat, R, T.gagg, st);
  PairBwdArgs pa{};
  pa.N = N; pa.L = L; pa.Lp = w.Lp; pa.z = z; pa.alpha = w.alpha; pa.mask = mask; pa.gfeat = T.gfeat; pa.g_agg = T.gagg; pa.Pm = T.Pm;
  pa.PG = T.PG; pa.R = R; pa.Wb = bw.Wb_raw; pa.coef = bw.coef; pa.g_log = T.glog; pa.G = T.G; pa.dz = dz;
  pa.dz_accumulate = dz_accumulate ? 1 : 0; pa.part = T.part;
  launch_pair_bwd(pa, st);
  bwd_colsum(M, 780, T.part, 780, nullptr, 0, T.tmp, T.scr, st, 1.f);
  CUDA_TRY(cudaMemcpyAsync(grad_of(m, p + "proj_pair_bias.weight"), T.tmp, H * C * 4, cudaMemcpyDeviceToDevice, st));
  coef_grad_kernel<<<1, 32, 0, st>>>(T.tmp + 768, bw.sc_raw, grad_of(m, p + "spatial_coef"));
  // ---- projections (ga.py:82-83,96-105,122,129-132)
  {
    // d x += G W_cat on the tensor cores: A = G (its lo plane goes where the plain projections were: they are dead after the pair
    // backward), B = W_cat^T [128][2016] hi | lo; the product lands in T.t1 and is added to gx
    float* wt_h = T.scr; float* wt_l = wt_h + (size_t)NPROJ * F;
    launch_lo(T.G, T.Pm, (size_t)M * NPROJ, st);
    launch_transpose_split(bw.Wcat, NPROJ, F, wt_h, wt_l, st);
    if (!launch_gemm3x_plain(M, F, NPROJ, T.G, T.Pm, NPROJ, wt_h, wt_l, NPROJ, T.t1, F, nullptr, st))
      return fail(ABOPT_ERR_CUDA, "cuTensorMapEncodeTiled failed (backward: d x)");
    bwd_add_mask(M, F, gx, T.t1, nullptr, gx, st);
  }
  bwd_wgrad(M, NPROJ, F, T.G, NPROJ, x, F, false, grad_of(m, p + "proj_query.weight"), T.scr, T.scr_floats, st);
  CHECK_LAUNCH();
  return ABOPT_OK;
}

extern "C" int abopt_ga_block_backward(abopt_model* m, int layer, int N, int L, const float* R, const float* t, const float* x,
                                       const float* z, const uint8_t* mask, const float* g_out, float* g_x, float* g_z, void* stream) {
  int rc = check_ready(m, N, L); if (rc) return rc;
  if (layer < 0 || layer >= m->cfg.num_layers) return fail(ABOPT_ERR_ARG, "layer out of range");
  if (!R || !t || !x || !z || !mask || !g_out || !g_x || !g_z) return fail(ABOPT_ERR_ARG, "null tensor");
  DeviceGuard g(m->device);
  rc = ensure_workspace(m, N, L); if (rc) return rc;
  TrainWS T;
  rc = ensure_train_ws(m, N, L, m->ws.Lp, T); if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemcpyAsync(T.gx, g_out, (size_t)N * L * F * 4, cudaMemcpyDeviceToDevice, st));
  rc = block_backward(m, T, layer, N, L, R, t, x, z, mask, T.gx, g_z, false, st); if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(g_x, T.gx, (size_t)N * L * F * 4, cudaMemcpyDeviceToDevice, st));
  return ABOPT_OK;
}

extern "C" int abopt_model_get_grad(abopt_model* m, const char* key, float* dst, size_t numel, void* stream) {
  if (!m || !key || !dst) return fail(ABOPT_ERR_ARG, "null argument");
  auto it = m->grad_off.find(key);
  if (it == m->grad_off.end()) return fail(ABOPT_ERR_KEY, std::string("no gradient for key: ") + key);
  if (it->second.second != numel) return fail(ABOPT_ERR_KEY, std::string("size mismatch for ") + key);
  DeviceGuard g(m->device);
  CUDA_TRY(cudaMemcpyAsync(dst, m->grad_buf + it->second.first, numel * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return ABOPT_OK;
}

// one 3-layer head (dpm_full.py:45-62; PerResiduePredictor, common/nn.py:164-188): recompute its activations from `in`
// (M, 131), backpropagate g (M, nout; row pitch ldg) and add d loss / d in to g_in (M, 131)
static void head_backward(abopt_model* m, TrainWS& T, int M, const abopt_model::RawHead& rh, const std::string& k0, const std::string& k2,
                          const std::string& k4, const float* in, const float* g, int ldg, float* g_in, float* out_raw, cudaStream_t st) {
  const int K0 = F + 3;
  bwd_gemm_nt(M, F, K0, in, K0, false, rh.W0, K0, rh.b0, T.a0, F, st);
  bwd_gemm_nt(M, F, F, T.a0, F, true, rh.W2, F, rh.b2, T.a1, F, st);
  if (out_raw) { bwd_gemm_nt(M, rh.nout, F, T.a1, F, true, rh.W4, F, rh.b4, out_raw, rh.nout, st); return; }
  bwd_wgrad(M, rh.nout, F, g, ldg, T.a1, F, true, grad_of(m, k4 + ".weight"), T.scr, T.scr_floats, st);
  bwd_colsum(M, rh.nout, g, ldg, nullptr, 0, grad_of(m, k4 + ".bias"), T.scr, st, 1.f);
  bwd_gemm_nn(M, rh.nout, F, g, ldg, rh.W4, F, T.t1, F, false, st);
  bwd_relu((size_t)M * F, T.t1, T.a1, st);
  bwd_wgrad(M, F, F, T.t1, F, T.a0, F, true, grad_of(m, k2 + ".weight"), T.scr, T.scr_floats, st);
  bwd_colsum(M, F, T.t1, F, nullptr, 0, grad_of(m, k2 + ".bias"), T.scr, st, 1.f);
  bwd_gemm_nn(M, F, F, T.t1, F, rh.W2, F, T.t2, F, false, st);
  bwd_relu((size_t)M * F, T.t2, T.a0, st);
  bwd_wgrad(M, F, K0, T.t2, F, in, K0, false, grad_of(m, k0 + ".weight"), T.scr, T.scr_floats, st);
  bwd_colsum(M, F, T.t2, F, nullptr, 0, grad_of(m, k0 + ".bias"), T.scr, st, 1.f);
  bwd_gemm_nn(M, F, K0, T.t2, F, rh.W0, K0, g_in, K0, true, st);
}

extern "C" int abopt_loss_backward(abopt_model* m, int N, int L, const float* v_0, const float* p_0, const int64_t* s_0,
                                   const float* res_feat, const float* pair_feat, const uint8_t* mask_generate,
                                   const uint8_t* mask_res, uint32_t flags, const int64_t* t, uint64_t seed,
                                   const abopt_step_noise* noise, const float* loss_weights, float* losses_out, float* d_res_feat,
                                   float* d_pair_feat, void* stream) {
  int rc = check_ready(m, N, L, ABOPT_SCOPE_FULL); if (rc) return rc;
  if (!v_0 || !p_0 || !s_0 || !res_feat || !pair_feat || !mask_generate || !mask_res || !t || !losses_out || !d_res_feat || !d_pair_feat)
    return fail(ABOPT_ERR_ARG, "null tensor");
  if (noise && (!noise->u || !noise->expo_ang || !noise->unif_ang || !noise->gauss_ang || !noise->z_pos || !noise->expo_seq))
    return fail(ABOPT_ERR_ARG, "incomplete abopt_step_noise record");
  DeviceGuard g(m->device);
  cudaStream_t st = (cudaStream_t)stream;
  rc = ensure_workspace(m, N, L); if (rc) return rc;
  Workspace& w = m->ws;
  TrainWS T;
  rc = ensure_train_ws(m, N, L, w.Lp, T); if (rc) return rc;
  const size_t M = (size_t)N * L;
  const int nl = m->cfg.num_layers;
  const size_t need = M * (3 + 3 + 3 + 6) * sizeof(float) + M * sizeof(long long) + (size_t)N * sizeof(float) + 256;
  if (m->train_bytes < need) {
    if (m->train_buf) { CUDA_TRY(cudaFree(m->train_buf)); m->train_buf = nullptr; m->train_bytes = 0; }
    CUDA_TRY(cudaMalloc(&m->train_buf, need));
    m->train_bytes = need;
  }
  long long* s_noisy = reinterpret_cast<long long*>(m->train_buf);
  float* v_noisy = reinterpret_cast<float*>(s_noisy + M);
  float* p_noisy = v_noisy + M * 3;
  float* zbuf = p_noisy + M * 3;
  float* rows = zbuf + M * 3;
  float* beta = rows + M * 6;
  const bool ds = (flags & ABOPT_SAMPLE_STRUCTURE) != 0, dq = (flags & ABOPT_SAMPLE_SEQUENCE) != 0;
  // ---- forward (dpm_full.py:156-234), evaluated as with autograd enabled: log_rotation clamps at -0.999 (so3.py:12-17)
  InitArgs ia{};
  ia.M = (int)M; ia.L = L; ia.T0 = 0;
  ia.sample_structure = ds ? 1 : 0; ia.sample_sequence = dq ? 1 : 0; ia.optimize = 1; ia.has_prmsd = 0;
  ia.v = v_0; ia.p_ang = p_0; ia.s = (const long long*)s_0; ia.mask_gen = mask_generate;
  ia.v_out = v_noisy; ia.p_out_ang = p_noisy; ia.s_out = s_noisy;
  ia.seed = seed; ia.row0 = (uint32_t)(m->batch_offset * L); ia.tvec = (const long long*)t; ia.seq_all_rows = 1; ia.z_out = ds ? zbuf : nullptr;
  ia.grad_clamp = 1;
  if (noise) {
    if (ds) launch_angle_argmax((int)M, L, (const long long*)t, 0, m->diff.ang_Y[0], noise->expo_ang, mask_generate, w.bin_idx, st);
    ia.add = NoisePtrs{noise->u, noise->unif_ang, noise->gauss_ang, noise->z_pos, noise->expo_seq, w.bin_idx};
  }
  launch_init(ia, m->diff, st);
  launch_gather_beta(N, (const long long*)t, m->diff.betas, beta, st);
  launch_mixer((int)M, res_feat, s_noisy, v_noisy, m->eps, T.xs, w.Rbuf, p_noisy, w.pnorm, m->diff.pos_mean, m->diff.pos_scale, nullptr, st);
  for (int l = 0; l < nl; ++l) {                                  // only the block INPUTS are kept
    rc = run_block(m, l, N, L, w.Rbuf, w.pnorm, T.xs + (size_t)l * M * F, nullptr, pair_feat, mask_res, T.xs + (size_t)(l + 1) * M * F, nullptr,
                   nullptr, st);
    if (rc) return rc;
  }
  const float* xL = T.xs + (size_t)nl * M * F;
  launch_heads((int)M, L, xL, beta, 1, w.Rbuf, v_noisy, mask_generate, m->eps, w.v_net, w.R_next, w.eps_pos, w.c_den, w.prmsd_rows,
               m->cfg.has_prmsd ? w.prmsd_logits : nullptr, st);
  LossArgs la{};
  la.N = N; la.L = L; la.abdock = m->cfg.has_prmsd ? 1 : 0; la.pred_x0 = m->cfg.obj_pred_x0 ? 1 : 0;
  la.has_prmsd = m->cfg.has_prmsd; la.bins = m->cfg.prmsd_bins; la.dmin = m->cfg.prmsd_min; la.dmax = m->cfg.prmsd_max;
  la.v_0 = v_0; la.p_0_ang = p_0; la.s_0 = (const long long*)s_0;
  la.p_noisy_ang = p_noisy; la.s_noisy = s_noisy; la.z = ds ? zbuf : nullptr;
  la.R_pred = w.R_next; la.p_pred = w.eps_pos; la.c_den = w.c_den; la.prmsd_logits = w.prmsd_logits;
  la.mask_gen = mask_generate; la.mask_res = mask_res; la.tvec = (const long long*)t;
  la.rows = rows; la.out = losses_out;
  launch_loss(la, m->diff, st);
  // ---- backward: losses -> heads
  const int K0 = F + 3;
  bwd_heads_cat((int)M, L, xL, beta, T.hcat, st);
  CUDA_TRY(cudaMemsetAsync(T.ghcat, 0, M * K0 * 4, st));
  head_backward(m, T, (int)M, m->raw.head[1], "", "", "", T.hcat, nullptr, 0, nullptr, T.orot, st);      // raw eps_rot_net output (M, 3)
  LossBwdArgs lb{};
  lb.N = N; lb.L = L; lb.abdock = la.abdock; lb.pred_x0 = la.pred_x0; lb.has_prmsd = la.has_prmsd; lb.bins = m->cfg.has_prmsd ? m->cfg.prmsd_bins : 1;
  lb.dmin = la.dmin; lb.dmax = la.dmax;
  for (int k = 0; k < 5; ++k) lb.lw[k] = loss_weights ? loss_weights[k] : 1.f;
  lb.v_0 = v_0; lb.p_0_ang = p_0; lb.s_0 = (const long long*)s_0; lb.p_noisy_ang = p_noisy; lb.s_noisy = s_noisy; lb.z = la.z;
  lb.R = w.Rbuf; lb.R_pred = w.R_next; lb.eps_pos = w.eps_pos; lb.c_den = w.c_den; lb.o_rot = T.orot; lb.ld_orot = 3;
  lb.prmsd_logits = w.prmsd_logits; lb.mask_gen = mask_generate; lb.mask_res = mask_res; lb.tvec = (const long long*)t; lb.rows = rows;
  lb.stats = T.stats; lb.glog = T.glogit; lb.GO = T.GO; lb.GP = m->cfg.has_prmsd ? T.GP : nullptr;
  launch_loss_bwd(lb, m->diff, st);
  const char* hn[3] = {"eps_net.eps_crd_net.", "eps_net.eps_rot_net.", "eps_net.eps_seq_net."};
  const int hoff[3] = {0, 3, 6};
  for (int hh = 0; hh < 3; ++hh)
    head_backward(m, T, (int)M, m->raw.head[hh], std::string(hn[hh]) + "0", std::string(hn[hh]) + "2", std::string(hn[hh]) + "4", T.hcat,
                  T.GO + hoff[hh], 26, T.ghcat, nullptr, st);
  if (m->cfg.has_prmsd) {                                          // pRMSD head: LayerNorm(131) -> three linears, mean over all L rows
    const std::string p = "eps_net.prmsd_predictor.";
    bwd_ln131_fwd((int)M, T.hcat, m->raw.prm_g, m->raw.prm_b, T.lnb, st);
    CUDA_TRY(cudaMemsetAsync(T.gcat, 0, M * K0 * 4, st));           // gcat doubles as d loss / d LayerNorm output here
    head_backward(m, T, (int)M, m->raw.head[3], p + "linear_1", p + "linear_2", p + "linear_3", T.lnb, T.GP, lb.bins, T.gcat, nullptr, st);
    bwd_colsum((int)M, K0, T.gcat, K0, nullptr, 0, grad_of(m, p + "layer_norm.beta"), T.scr, st, 1.f);
    bwd_ln131_bwd((int)M, T.gcat, T.hcat, m->raw.prm_g, T.ghcat, T.lnb, st);
    bwd_colsum((int)M, K0, T.lnb, K0, nullptr, 0, grad_of(m, p + "layer_norm.gamma"), T.scr, st, 1.f);
  }
  strided_copy_kernel<<<(unsigned)((M * F + 255) / 256), 256, 0, st>>>((int)M, F, T.ghcat, K0, T.gx, F);
  // ---- encoder, last block first
  for (int l = nl - 1; l >= 0; --l) {
    rc = block_backward(m, T, l, N, L, w.Rbuf, w.pnorm, T.xs + (size_t)l * M * F, pair_feat, mask_res, T.gx, d_pair_feat, l != nl - 1, st);
    if (rc) return rc;
  }
  // ---- mixer and sequence embedding (dpm_full.py:86-88)
  bwd_mixer_cat((int)M, res_feat, s_noisy, m->eps.emb, T.cat0, st);
  bwd_gemm_nt((int)M, F, 2 * F, T.cat0, 2 * F, false, m->raw.Wm0, 2 * F, m->raw.bm0, T.a0, F, st);          // m_a
  bwd_wgrad((int)M, F, F, T.gx, F, T.a0, F, true, grad_of(m, "eps_net.res_feat_mixer.2.weight"), T.scr, T.scr_floats, st);
  bwd_colsum((int)M, F, T.gx, F, nullptr, 0, grad_of(m, "eps_net.res_feat_mixer.2.bias"), T.scr, st, 1.f);
  bwd_gemm_nn((int)M, F, F, T.gx, F, m->raw.Wm2, F, T.t1, F, false, st);
  bwd_relu(M * F, T.t1, T.a0, st);
  bwd_wgrad((int)M, F, 2 * F, T.t1, F, T.cat0, 2 * F, false, grad_of(m, "eps_net.res_feat_mixer.0.weight"), T.scr, T.scr_floats, st);
  bwd_colsum((int)M, F, T.t1, F, nullptr, 0, grad_of(m, "eps_net.res_feat_mixer.0.bias"), T.scr, st, 1.f);
  bwd_gemm_nn((int)M, F, 2 * F, T.t1, F, m->raw.Wm0, 2 * F, T.gcat, 2 * F, false, st);
  strided_copy_kernel<<<(unsigned)((M * F + 255) / 256), 256, 0, st>>>((int)M, F, T.gcat, 2 * F, d_res_feat, F);
  bwd_embed_grad((int)M, s_noisy, T.gcat, grad_of(m, "eps_net.current_sequence_embedding.weight"), st);
  CHECK_LAUNCH();
  return ABOPT_OK;
}
