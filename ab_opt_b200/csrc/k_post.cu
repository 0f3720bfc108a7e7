// The steps right after the sampling loop, stateless device entry points:
//   abopt_reconstruct_backbone_partially   reconstruct_backbone_partially, /root/reference/AbDock/src/modules/common/geometry.py:450-480
//                                          (reconstruct_backbone :404-447, local_to_global :72-92, compose_chain :120-140,
//                                           get_backbone_dihedral_angles :307-348, topology.py:5-24)
//   abopt_pairwise_rmsd                    calc_per_rmsd / calc_avg_rmsd, AbDock/src/tools/runner/design_for_testset.py:556-570
//   abopt_rank_commoness                   rank_commoness, design_for_testset.py:573-589
// The reference runs the first on CPU tensors once per trajectory frame (101 times per complex when traj.pdb is written,
// tools/runner/design_for_pdb.py:166-209) and the others on up to 1000 candidates; here the frames of a whole trajectory go
// through ONE launch (the leading dimension N is frames x complexes).
#include <algorithm>
#include <cmath>
#include <string>

#include "../../include/abopt_b200.h"
#include "kernels.h"

namespace abopt {
int api_fail(int code, const std::string& msg);

namespace {
// q = R p + t (local_to_global, geometry.py:72-92); R row-major
__device__ __forceinline__ void to_global(const float* R, const float* t, const float* p, float* q) {
#pragma unroll
  for (int i = 0; i < 3; ++i) q[i] = R[i * 3] * p[0] + R[i * 3 + 1] * p[1] + R[i * 3 + 2] * p[2] + t[i];
}

struct ReconArgs {
  long long rows; int L, A;
  const float* pos_ctx; const float* R; const float* t; const long long* aa; const long long* chain_nb; const long long* res_nb;
  const uint8_t* mask_atoms; const uint8_t* mask_recons; const float* bb; const float* ox;
  float* pos_new; uint8_t* mask_new;
};

// one thread per residue
__global__ void __launch_bounds__(128) reconstruct_kernel(ReconArgs a) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= a.rows) return;
  const int A = a.A;
  float* out = a.pos_new + (size_t)row * A * 3;
  uint8_t* mout = a.mask_new + (size_t)row * A;
  if (!a.mask_recons[row]) {                                                          // geometry.py:467-477: context residues pass through
    const float* src = a.pos_ctx + (size_t)row * A * 3;
    for (int k = 0; k < A * 3; ++k) out[k] = src[k];
    for (int k = 0; k < A; ++k) mout[k] = a.mask_atoms[(size_t)row * A + k];
    return;
  }
  long long s = a.aa[row];
  const int aa = (int)(s < 0 ? 0 : (s > 20 ? 20 : s));                                // geometry.py:419
  const float* R = a.R + (size_t)row * 9;
  const float* t = a.t + (size_t)row * 3;
  float q[4][3];
#pragma unroll
  for (int k = 0; k < 3; ++k) to_global(R, t, a.bb + (aa * 3 + k) * 3, q[k]);         // N, CA, C (geometry.py:421-423)
  // psi needs the N of the next residue when the two are bonded (geometry.py:340-343, topology.py:13-16)
  const int i = (int)(row % a.L);
  float psi = 0.f;
  if (i < a.L - 1) {
    long long d = a.res_nb[row + 1] - a.res_nb[row];
    d = d < 0 ? -d : d;
    const bool bonded = d == 1 && a.chain_nb[row + 1] == a.chain_nb[row] && a.mask_atoms[(size_t)row * A + 1];
    long long s2 = a.aa[row + 1];
    const int aa2 = (int)(s2 < 0 ? 0 : (s2 > 20 ? 20 : s2));
    float n_next[3];
    to_global(a.R + (size_t)(row + 1) * 9, a.t + (size_t)(row + 1) * 3, a.bb + (aa2 * 3) * 3, n_next);
    const float x = dihedral4(q[0], q[1], q[2], n_next);
    psi = bonded ? x : 0.f;
  }
  // O = R Rx(psi) o + t (geometry.py:427-444)
  float sn, cs;
  sincosf(psi, &sn, &cs);
  const float* o = a.ox + aa * 3;
  const float turned[3] = {o[0], cs * o[1] - sn * o[2], sn * o[1] + cs * o[2]};
  to_global(R, t, turned, q[3]);
#pragma unroll
  for (int k = 0; k < 4; ++k) { out[k * 3] = q[k][0]; out[k * 3 + 1] = q[k][1]; out[k * 3 + 2] = q[k][2]; }
  for (int k = 12; k < A * 3; ++k) out[k] = 0.f;                                      // F.pad(..., value=0), geometry.py:465
  for (int k = 0; k < A; ++k) mout[k] = k < 4 ? 1 : 0;
}

// rmsd[i][j] = sqrt(mean_m |s_i[m] - s_j[m]|^2); one CTA per structure i, warp per j (strided), row sums in double
__global__ void __launch_bounds__(256) rmsd_rows_kernel(int B, int M, const float* __restrict__ S, float* __restrict__ rmsd, float* __restrict__ score) {
  extern __shared__ float sI[];                      // structure i, M * 3 floats
  __shared__ double sPart[8];
  const int i = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < M * 3; k += blockDim.x) sI[k] = S[(size_t)i * M * 3 + k];
  __syncthreads();
  double rowsum = 0.0;
  for (int j = wid; j < B; j += 8) {
    const float* sj = S + (size_t)j * M * 3;
    float acc = 0.f;
    for (int m = lane; m < M; m += 32) {
      const float dx = sI[m * 3] - sj[m * 3], dy = sI[m * 3 + 1] - sj[m * 3 + 1], dz = sI[m * 3 + 2] - sj[m * 3 + 2];
      acc += dx * dx + dy * dy + dz * dz;
    }
    acc = warp_sum(acc);
    const float r = sqrtf(acc / (float)M);
    if (lane == 0) {
      if (rmsd) rmsd[(size_t)i * B + j] = r;
      rowsum += (double)r;
    }
  }
  if (lane == 0) sPart[wid] = rowsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int k = 0; k < 8; ++k) tot += sPart[k];
    score[i] = (float)(tot / (double)(B - 1));       // design_for_testset.py:586
  }
}

// avg = sum_i score_i / B  (= rmsd.sum() / (B (B - 1)), design_for_testset.py:569); rank by counting, ties by index
__global__ void __launch_bounds__(256) rank_kernel(int B, int k, const float* __restrict__ score, long long* __restrict__ rank, float* __restrict__ avg) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
    const float v = score[i];
    int pos = 0;
    for (int j = 0; j < B; ++j) {
      const float u = score[j];
      pos += (u < v || (u == v && j < i)) ? 1 : 0;
    }
    if (rank && pos < k) rank[pos] = i;
  }
  if (avg && blockIdx.x == 0 && threadIdx.x == 0) {
    double tot = 0.0;
    for (int j = 0; j < B; ++j) tot += (double)score[j];
    *avg = (float)(tot / (double)B);
  }
}
}  // namespace
}  // namespace abopt

using namespace abopt;

static int post_device_check() {
  int dev = 0, major = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);     // cheap: these run per call
  if (e != cudaSuccess) return api_fail(ABOPT_ERR_CUDA, std::string("no usable CUDA device: ") + cudaGetErrorString(e));
  if (major != 10) return api_fail(ABOPT_ERR_CUDA, "libabopt_b200 needs a B200-class GPU (sm_100); the current device has compute capability " +
                                                       std::to_string(major) + ".x");
  return ABOPT_OK;
}

extern "C" int abopt_reconstruct_backbone_partially(int N, int L, int A, const float* pos_ctx, const float* R_new, const float* t_new,
                                                    const int64_t* aa, const int64_t* chain_nb, const int64_t* res_nb,
                                                    const uint8_t* mask_atoms, const uint8_t* mask_recons, const float* bb_table,
                                                    const float* o_table, float* pos_new, uint8_t* mask_new, void* stream) {
  if (N < 0 || L < 0) return api_fail(ABOPT_ERR_ARG, "negative size");
  if (A < 4) return api_fail(ABOPT_ERR_ARG, "at least 4 atoms per residue (N, CA, C, O) are needed");
  if (N == 0 || L == 0) return ABOPT_OK;
  if (!pos_ctx || !R_new || !t_new || !aa || !chain_nb || !res_nb || !mask_atoms || !mask_recons || !bb_table || !o_table || !pos_new || !mask_new)
    return api_fail(ABOPT_ERR_ARG, "null tensor");
  if (int rc = post_device_check()) return rc;
  ReconArgs a{(long long)N * L, L, A, pos_ctx, R_new, t_new, (const long long*)aa, (const long long*)chain_nb, (const long long*)res_nb,
              mask_atoms, mask_recons, bb_table, o_table, pos_new, mask_new};
  cudaStream_t st = (cudaStream_t)stream;
  {
    ProfScope ps(KK_OTHER, st);
    reconstruct_kernel<<<(unsigned)((a.rows + 127) / 128), 128, 0, st>>>(a);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return api_fail(ABOPT_ERR_CUDA, std::string("reconstruct_kernel: ") + cudaGetErrorString(e));
  return ABOPT_OK;
}

static int rmsd_common(int B, int M, const float* structures, float* rmsd, float* score, int k, int64_t* rank, float* avg, void* stream) {
  if (B < 2 || M < 1) return api_fail(ABOPT_ERR_ARG, "need at least two structures of at least one point");
  if ((size_t)M * 3 * sizeof(float) > 160 * 1024) return api_fail(ABOPT_ERR_ARG, "structures longer than 13653 points are not supported");
  if (!structures || !score) return api_fail(ABOPT_ERR_ARG, "null tensor");
  if (rank && (k < 1 || k > B)) return api_fail(ABOPT_ERR_ARG, "k out of range");
  if (int rc = post_device_check()) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)M * 3 * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(rmsd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return api_fail(ABOPT_ERR_CUDA, std::string("rmsd_rows_kernel shared memory: ") + cudaGetErrorString(e));
  }
  {
    ProfScope ps(KK_OTHER, st);
    rmsd_rows_kernel<<<B, 256, smem, st>>>(B, M, structures, rmsd, score);
  }
  if (rank || avg) {
    ProfScope ps(KK_OTHER, st);
    rank_kernel<<<rank ? std::min((B + 255) / 256, 148) : 1, 256, 0, st>>>(B, k, score, (long long*)rank, avg);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return api_fail(ABOPT_ERR_CUDA, std::string("rmsd kernels: ") + cudaGetErrorString(e));
  return ABOPT_OK;
}

extern "C" int abopt_pairwise_rmsd(int B, int M, const float* structures, float* rmsd, float* score, float* avg, void* stream) {
  return rmsd_common(B, M, structures, rmsd, score, 0, nullptr, avg, stream);
}

extern "C" int abopt_rank_commoness(int B, int M, const float* structures, int k, float* score, int64_t* rank, void* stream) {
  if (!rank) return api_fail(ABOPT_ERR_ARG, "null tensor");
  return rmsd_common(B, M, structures, nullptr, score, k, rank, nullptr, stream);
}
