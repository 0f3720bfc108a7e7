// Per-residue features (second half): ResidueEmbedding.forward,
// /root/reference/AbDock/src/modules/encoders/residue.py:27-94 (AbDesign: diffab/modules/encoders/residue.py), with
// construct_3d_basis (modules/common/geometry.py:47-69), global_to_local (:94-113), get_backbone_dihedral_angles (:307-348),
// get_terminus_flag / get_consecutive_flag (modules/common/topology.py:5-24) and AngularEncoding (layers.py:85-106).
//
// One kernel, four residues per CTA pass.  The reference scatters the local atom coordinates into a (22 x A x 3) one-hot slot
// and multiplies by a 1285-column weight; here the slot selects the 3A weight rows that matter, and the amino-acid and fragment
// type embeddings enter the first layer as pre-multiplied tables (T = E W^T), so a residue costs 84 + 256 + 128 + 128 MACs per
// output instead of 1285 + ...  B x L is 16k rows: this kernel is launch- and latency-bound (well under a millisecond), the
// O(L^2) work of the featurisation is pair_embed_kernel's.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/abopt_b200.h"
#include "kernels.h"

namespace abopt {
int api_fail(int code, const std::string& msg);

namespace {
constexpr int RE_THREADS = 256;
constexpr int RE_RT = 4;             // residues per CTA pass
constexpr int RE_F = 128;
constexpr int RE_AA = 22, RE_UNK = 20, RE_MAXA = 15, RE_TYPES = 10, RE_ANG = 39;

struct ResEmbedW {
  int A;
  const float* Taa;      // [22][256]   aatype_embed . W1[:, 0:128]^T
  const float* Ttype;    // [10][256]   type_embed   . W1[:, 1157 + ...]^T
  const float* W1c;      // [22 * A * 3][256]  coordinate columns of mlp.0, transposed
  const float* W1d;      // [39][256]          dihedral columns of mlp.0, transposed
  const float* W2;       // [256][128]  mlp.2^T
  const float* W3;       // [128][128]  mlp.4^T
  const float* W4;       // [128][128]  mlp.6^T
  const float* b1; const float* b2; const float* b3; const float* b4;
  float freq[6];
};
struct ResEmbedArgs {
  int N, L, A_in;
  const long long* aa; const long long* res_nb; const long long* chain_nb; const long long* fragment_type;
  const float* pos; const uint8_t* mask_atoms; const uint8_t* structure_mask; const uint8_t* sequence_mask;
  float* out;
};

__global__ void __launch_bounds__(RE_THREADS) res_embed_kernel(ResEmbedW w, ResEmbedArgs a) {
  __shared__ float sCrd[RE_RT][RE_MAXA * 3 + 3];
  __shared__ float sAng[RE_RT][RE_ANG + 1];
  __shared__ float sH1[RE_RT][256];
  __shared__ float sH2[RE_RT][RE_F];
  __shared__ float sH3[RE_RT][RE_F];
  __shared__ int sAa[RE_RT], sType[RE_RT], sOk[RE_RT];
  const int tid = threadIdx.x, A = w.A, L = a.L, A_in = a.A_in;
  const long long rows = (long long)a.N * L;
  const long long groups = (rows + RE_RT - 1) / RE_RT;
  for (long long g = blockIdx.x; g < groups; g += gridDim.x) {
    const long long row0 = g * RE_RT;
    __syncthreads();
    // ---- features
    if (tid < RE_RT * RE_MAXA) {                                     // local coordinates of one atom (residue.py:52-61)
      const int r = tid / RE_MAXA, at = tid - r * RE_MAXA;
      const long long row = row0 + r;
      float o[3] = {0.f, 0.f, 0.f};
      if (row < rows && at < A) {
        const float* P = a.pos + (size_t)row * A_in * 3;
        const float ca[3] = {P[3], P[4], P[5]};
        float e1[3], v2[3], e2[3], e3[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { e1[c] = P[6 + c] - ca[c]; v2[c] = P[c] - ca[c]; }             // C - CA, N - CA
        const float n1 = sqrtf(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]) + 1e-6f;               // geometry.py:32-33, 57-58
#pragma unroll
        for (int c = 0; c < 3; ++c) e1[c] = e1[c] / n1;
        const float pr = e1[0] * v2[0] + e1[1] * v2[1] + e1[2] * v2[2];                              // geometry.py:44, 61
#pragma unroll
        for (int c = 0; c < 3; ++c) e2[c] = v2[c] - pr * e1[c];
        const float n2 = sqrtf(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]) + 1e-6f;
#pragma unroll
        for (int c = 0; c < 3; ++c) e2[c] = e2[c] / n2;
        e3[0] = e1[1] * e2[2] - e1[2] * e2[1]; e3[1] = e1[2] * e2[0] - e1[0] * e2[2]; e3[2] = e1[0] * e2[1] - e1[1] * e2[0];
        const bool keep = a.mask_atoms[(size_t)row * A_in + at] && (!a.structure_mask || a.structure_mask[row]);   // residue.py:60-61, 69-71
        if (keep) {
          const float q[3] = {P[at * 3] - ca[0], P[at * 3 + 1] - ca[1], P[at * 3 + 2] - ca[2]};
          o[0] = e1[0] * q[0] + e1[1] * q[1] + e1[2] * q[2];                                         // R^T (x - t), geometry.py:110-112
          o[1] = e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2];
          o[2] = e3[0] * q[0] + e3[1] * q[1] + e3[2] * q[2];
        }
      }
      sCrd[r][at * 3] = o[0]; sCrd[r][at * 3 + 1] = o[1]; sCrd[r][at * 3 + 2] = o[2];
    } else if (tid >= 64 && tid < 64 + RE_RT * 3) {                  // one backbone dihedral (geometry.py:307-348)
      const int r = (tid - 64) / 3, k = (tid - 64) - r * 3;         // k: 0 omega, 1 phi, 2 psi
      const long long row = row0 + r;
      float x = 0.f;
      bool valid = false;
      if (row < rows) {
        const int i = (int)(row % L);
        const long long nb = (k < 2) ? row - 1 : row + 1;           // the bonded neighbour this angle needs
        const bool inside = (k < 2) ? (i > 0) : (i < L - 1);
        if (inside) {
          const long long lo = (k < 2) ? nb : row;                  // bond lo -> lo + 1; topology.py:13-16 masks with mask[lo]
          long long d = a.res_nb[lo + 1] - a.res_nb[lo];
          d = d < 0 ? -d : d;
          valid = d == 1 && a.chain_nb[lo + 1] == a.chain_nb[lo] && a.mask_atoms[(size_t)lo * A_in + 1];
          const float* Pa = a.pos + (size_t)lo * A_in * 3;          // residue lo:     N 0..2, CA 3..5, C 6..8
          const float* Pb = Pa + (size_t)A_in * 3;                  // residue lo + 1
          x = k == 0 ? dihedral4(Pa + 3, Pa + 6, Pb, Pb + 3) : (k == 1 ? dihedral4(Pa + 6, Pb, Pb + 3, Pb + 6) : dihedral4(Pa, Pa + 3, Pa + 6, Pb));
        }
        if (valid && a.structure_mask) {                            // residue.py:77-86: the residue and both (rolled) neighbours
          const long long base = row - i;
          const int im = i == 0 ? L - 1 : i - 1, ip = i == L - 1 ? 0 : i + 1;
          valid = a.structure_mask[row] && a.structure_mask[base + im] && a.structure_mask[base + ip];
        }
      }
      const float s = valid ? 1.f : 0.f;
      x = valid ? x : 0.f;
      float* dst = &sAng[r][k * 13];
      dst[0] = x * s;
#pragma unroll
      for (int f = 0; f < 6; ++f) {
        float sn, cs;
        sincosf(x * w.freq[f], &sn, &cs);
        dst[1 + f] = sn * s;
        dst[7 + f] = cs * s;
      }
    } else if (tid >= 96 && tid < 96 + RE_RT) {
      const int r = tid - 96;
      const long long row = row0 + r;
      int aa = 0, ft = 0, ok = 0;
      if (row < rows) {
        long long v = a.aa[row];
        if (a.sequence_mask && !a.sequence_mask[row]) v = RE_UNK;                                    // residue.py:47-49
        aa = (int)(v < 0 ? 0 : (v >= RE_AA ? RE_AA - 1 : v));
        v = a.fragment_type[row];
        ft = (int)(v < 0 ? 0 : (v >= RE_TYPES ? RE_TYPES - 1 : v));
        ok = a.mask_atoms[(size_t)row * A_in + 1] ? 1 : 0;                                           // residue.py:40, 93
      }
      sAa[r] = aa; sType[r] = ft; sOk[r] = ok;
    }
    __syncthreads();
    // ---- layer 1: 256 outputs, one per thread, four residues each
    {
      const int o = tid;
      float acc[RE_RT];
      const float b = w.b1[o];
#pragma unroll
      for (int r = 0; r < RE_RT; ++r) acc[r] = b + w.Taa[sAa[r] * 256 + o] + w.Ttype[sType[r] * 256 + o];
#pragma unroll 3
      for (int k = 0; k < RE_ANG; ++k) {
        const float wk = w.W1d[k * 256 + o];
#pragma unroll
        for (int r = 0; r < RE_RT; ++r) acc[r] = fmaf(sAng[r][k], wk, acc[r]);
      }
      const int K = A * 3;
#pragma unroll
      for (int r = 0; r < RE_RT; ++r) {
        const float* Wc = w.W1c + (size_t)sAa[r] * K * 256 + o;
        float s = 0.f;
#pragma unroll 5
        for (int k = 0; k < K; ++k) s = fmaf(sCrd[r][k], Wc[k * 256], s);
        acc[r] += s;
      }
#pragma unroll
      for (int r = 0; r < RE_RT; ++r) sH1[r][o] = fmaxf(acc[r], 0.f);
    }
    __syncthreads();
    // ---- layers 2..4: 128 outputs; thread = (output, residue pair)
    const int o = tid & 127, r0 = (tid >> 7) * 2;
    {
      float a0 = w.b2[o], a1 = a0;
#pragma unroll 8
      for (int k = 0; k < 256; ++k) {
        const float wk = w.W2[k * RE_F + o];
        a0 = fmaf(sH1[r0][k], wk, a0); a1 = fmaf(sH1[r0 + 1][k], wk, a1);
      }
      sH2[r0][o] = fmaxf(a0, 0.f); sH2[r0 + 1][o] = fmaxf(a1, 0.f);
    }
    __syncthreads();
    {
      float a0 = w.b3[o], a1 = a0;
#pragma unroll 8
      for (int k = 0; k < RE_F; ++k) {
        const float wk = w.W3[k * RE_F + o];
        a0 = fmaf(sH2[r0][k], wk, a0); a1 = fmaf(sH2[r0 + 1][k], wk, a1);
      }
      sH3[r0][o] = fmaxf(a0, 0.f); sH3[r0 + 1][o] = fmaxf(a1, 0.f);
    }
    __syncthreads();
    {
      float a0 = w.b4[o], a1 = a0;
#pragma unroll 8
      for (int k = 0; k < RE_F; ++k) {
        const float wk = w.W4[k * RE_F + o];
        a0 = fmaf(sH3[r0][k], wk, a0); a1 = fmaf(sH3[r0 + 1][k], wk, a1);
      }
      if (row0 + r0 < rows) a.out[(size_t)(row0 + r0) * RE_F + o] = sOk[r0] ? a0 : 0.f;             // residue.py:93
      if (row0 + r0 + 1 < rows) a.out[(size_t)(row0 + r0 + 1) * RE_F + o] = sOk[r0 + 1] ? a1 : 0.f;
    }
  }
}
}  // namespace
}  // namespace abopt

using namespace abopt;

// ------------------------------------------------------------------------------------------ C ABI
struct abopt_res_embed {
  int device = 0, A = 0;
  bool finalized = false;
  std::map<std::string, size_t> spec;
  std::map<std::string, std::vector<float>> sd;
  void* wbase = nullptr;
  ResEmbedW w;
  int sm_count = 148;
};

extern "C" int abopt_res_embed_create(int max_num_atoms, int device, abopt_res_embed** out) {
  if (!out) return api_fail(ABOPT_ERR_ARG, "null argument");
  if (max_num_atoms < 4 || max_num_atoms > RE_MAXA) return api_fail(ABOPT_ERR_ARG, "max_num_atoms must be in [4, 15] (backbone N, CA, C, O at least)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess) return api_fail(ABOPT_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return api_fail(ABOPT_ERR_ARG, "no such CUDA device");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return api_fail(ABOPT_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
  if (prop.major != 10) return api_fail(ABOPT_ERR_CUDA, std::string("libabopt_b200 needs a B200-class GPU (sm_100); found ") + prop.name);
  abopt_res_embed* re = new abopt_res_embed();
  re->device = device; re->A = max_num_atoms; re->sm_count = prop.multiProcessorCount;
  const size_t K1 = 2 * RE_F + (size_t)RE_AA * max_num_atoms * 3 + RE_ANG;
  re->spec = {{"aatype_embed.weight", (size_t)RE_AA * RE_F}, {"dihed_embed.freq_bands", 6}, {"type_embed.weight", (size_t)RE_TYPES * RE_F},
              {"mlp.0.weight", 256 * K1}, {"mlp.0.bias", 256}, {"mlp.2.weight", (size_t)RE_F * 256}, {"mlp.2.bias", RE_F},
              {"mlp.4.weight", (size_t)RE_F * RE_F}, {"mlp.4.bias", RE_F}, {"mlp.6.weight", (size_t)RE_F * RE_F}, {"mlp.6.bias", RE_F}};
  *out = re;
  return ABOPT_OK;
}

extern "C" void abopt_res_embed_destroy(abopt_res_embed* re) {
  if (!re) return;
  if (re->wbase) {
    int cur = 0;
    cudaGetDevice(&cur); cudaSetDevice(re->device);
    cudaFree(re->wbase);
    cudaSetDevice(cur);
  }
  delete re;
}

extern "C" int abopt_res_embed_set_tensor(abopt_res_embed* re, const char* key, const float* data, size_t numel, int on_device) {
  if (!re || !key) return api_fail(ABOPT_ERR_ARG, "null argument");
  auto it = re->spec.find(key);
  if (it == re->spec.end()) return api_fail(ABOPT_ERR_KEY, std::string("unexpected state-dict key: ") + key);
  if (it->second != numel) return api_fail(ABOPT_ERR_KEY, std::string("size mismatch for ") + key + ": expected " + std::to_string(it->second) +
                                                              " elements, got " + std::to_string(numel));
  if (!data) return api_fail(ABOPT_ERR_ARG, "null data");
  std::vector<float>& t = re->sd[key];
  t.resize(numel);
  if (on_device) {
    int cur = 0;
    cudaGetDevice(&cur); cudaSetDevice(re->device);
    cudaError_t e = cudaMemcpy(t.data(), data, numel * sizeof(float), cudaMemcpyDeviceToHost);
    cudaSetDevice(cur);
    if (e != cudaSuccess) return api_fail(ABOPT_ERR_CUDA, std::string("cudaMemcpy: ") + cudaGetErrorString(e));
  } else {
    memcpy(t.data(), data, numel * sizeof(float));
  }
  re->finalized = false;
  return ABOPT_OK;
}

extern "C" int abopt_res_embed_finalize(abopt_res_embed* re) {
  if (!re) return api_fail(ABOPT_ERR_ARG, "null argument");
  for (auto& kv : re->spec)
    if (!re->sd.count(kv.first)) return api_fail(ABOPT_ERR_STATE, "missing state-dict key: " + kv.first);
  const int A = re->A, KC = RE_AA * A * 3, K1 = 2 * RE_F + KC + RE_ANG;
  const std::vector<float>& W1 = re->sd["mlp.0.weight"];                                             // (256, K1): [aa 128 | crd 22*A*3 | dihed 39 | type 128]
  std::vector<float> img;
  auto reserve = [&](size_t n) { size_t off = (img.size() + 63) & ~size_t(63); img.resize(off + n, 0.f); return off; };
  const size_t o_taa = reserve(RE_AA * 256), o_tty = reserve(RE_TYPES * 256), o_w1c = reserve((size_t)KC * 256), o_w1d = reserve(RE_ANG * 256),
               o_w2 = reserve(256 * RE_F), o_w3 = reserve(RE_F * RE_F), o_w4 = reserve(RE_F * RE_F), o_b = reserve(256 + 3 * RE_F);
  auto premul = [&](const std::vector<float>& E, int rows, int col0, size_t off) {                  // T[r][o] = sum_c E[r][c] W1[o][col0 + c]
    for (int r = 0; r < rows; ++r)
      for (int o = 0; o < 256; ++o) {
        double s = 0.0;
        for (int c = 0; c < RE_F; ++c) s += (double)E[(size_t)r * RE_F + c] * (double)W1[(size_t)o * K1 + col0 + c];
        img[off + (size_t)r * 256 + o] = (float)s;
      }
  };
  premul(re->sd["aatype_embed.weight"], RE_AA, 0, o_taa);
  premul(re->sd["type_embed.weight"], RE_TYPES, RE_F + KC + RE_ANG, o_tty);
  auto transpose = [&](const float* Wsrc, int n_out, int ld, int col0, int K, size_t off) {          // dst[k][o] = W[o][col0 + k]
    for (int k = 0; k < K; ++k)
      for (int o = 0; o < n_out; ++o) img[off + (size_t)k * n_out + o] = Wsrc[(size_t)o * ld + col0 + k];
  };
  transpose(W1.data(), 256, K1, RE_F, KC, o_w1c);
  transpose(W1.data(), 256, K1, RE_F + KC, RE_ANG, o_w1d);
  transpose(re->sd["mlp.2.weight"].data(), RE_F, 256, 0, 256, o_w2);
  transpose(re->sd["mlp.4.weight"].data(), RE_F, RE_F, 0, RE_F, o_w3);
  transpose(re->sd["mlp.6.weight"].data(), RE_F, RE_F, 0, RE_F, o_w4);
  memcpy(&img[o_b], re->sd["mlp.0.bias"].data(), 256 * sizeof(float));
  memcpy(&img[o_b + 256], re->sd["mlp.2.bias"].data(), RE_F * sizeof(float));
  memcpy(&img[o_b + 256 + RE_F], re->sd["mlp.4.bias"].data(), RE_F * sizeof(float));
  memcpy(&img[o_b + 256 + 2 * RE_F], re->sd["mlp.6.bias"].data(), RE_F * sizeof(float));
  int cur = 0;
  cudaGetDevice(&cur); cudaSetDevice(re->device);
  if (re->wbase) { cudaFree(re->wbase); re->wbase = nullptr; }
  cudaError_t e = cudaMalloc(&re->wbase, img.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(re->wbase, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice);
  cudaSetDevice(cur);
  if (e != cudaSuccess) return api_fail(ABOPT_ERR_CUDA, std::string("residue-embed weights: ") + cudaGetErrorString(e));
  const float* base = static_cast<const float*>(re->wbase);
  re->w.A = A;
  re->w.Taa = base + o_taa; re->w.Ttype = base + o_tty; re->w.W1c = base + o_w1c; re->w.W1d = base + o_w1d;
  re->w.W2 = base + o_w2; re->w.W3 = base + o_w3; re->w.W4 = base + o_w4;
  re->w.b1 = base + o_b; re->w.b2 = base + o_b + 256; re->w.b3 = base + o_b + 256 + RE_F; re->w.b4 = base + o_b + 256 + 2 * RE_F;
  memcpy(re->w.freq, re->sd["dihed_embed.freq_bands"].data(), 6 * sizeof(float));
  re->finalized = true;
  return ABOPT_OK;
}

extern "C" int abopt_res_embed_forward(abopt_res_embed* re, int N, int L, int num_atoms_in, const int64_t* aa, const int64_t* res_nb,
                                       const int64_t* chain_nb, const float* pos_atoms, const uint8_t* mask_atoms,
                                       const int64_t* fragment_type, const uint8_t* structure_mask, const uint8_t* sequence_mask,
                                       float* res_feat, void* stream) {
  if (!re) return api_fail(ABOPT_ERR_ARG, "null handle");
  if (!re->finalized) return api_fail(ABOPT_ERR_STATE, "residue embedding not finalised");
  if (N < 0 || L < 0) return api_fail(ABOPT_ERR_ARG, "negative size");
  if (num_atoms_in < re->A) return api_fail(ABOPT_ERR_ARG, "pos_atoms / mask_atoms have fewer atoms per residue than max_num_atoms");
  if (N == 0 || L == 0) return ABOPT_OK;
  if (!aa || !res_nb || !chain_nb || !pos_atoms || !mask_atoms || !fragment_type || !res_feat) return api_fail(ABOPT_ERR_ARG, "null tensor");
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != re->device) cudaSetDevice(re->device);
  ResEmbedArgs a{N, L, num_atoms_in, (const long long*)aa, (const long long*)res_nb, (const long long*)chain_nb,
                 (const long long*)fragment_type, pos_atoms, mask_atoms, structure_mask, sequence_mask, res_feat};
  const long long groups = ((long long)N * L + RE_RT - 1) / RE_RT;
  const int grid = (int)std::min<long long>(groups, (long long)re->sm_count * 8);
  cudaStream_t st = (cudaStream_t)stream;
  {
    ProfScope ps(KK_OTHER, st);
    res_embed_kernel<<<grid, RE_THREADS, 0, st>>>(re->w, a);
  }
  cudaError_t e = cudaGetLastError();
  if (cur != re->device) cudaSetDevice(cur);
  if (e != cudaSuccess) return api_fail(ABOPT_ERR_CUDA, std::string("res_embed_kernel: ") + cudaGetErrorString(e));
  return ABOPT_OK;
}
