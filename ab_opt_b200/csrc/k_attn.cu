// Parity taps of the attention weights.  The pairwise part of a GABlock lives in
//   k_tc.cu       gemm3x_kernel<EpiProjPack>  projections -> packed attention operands
//   k_attn_tc.cu  attn_logits_tc_kernel       logits + softmax (tcgen05),   aggr_tc_kernel  node / point aggregation (tcgen05)
//   k_pair.cu     pair_bias_kernel (hoisted z . W_b),  pair_stream_kernel (streams z: pair aggregation)
#include "common.cuh"
#include "params.cuh"
#include "kernels.h"

namespace abopt {

// ------------------------------------------------------------------------------------------ taps
// alpha [chunk][h][i][Lp]  ->  reference layout (N, L, L, 12)   (parity taps only)
__global__ void alpha_to_reference_layout(int L, int Lp, int b0, const float* __restrict__ alpha, float* __restrict__ out) {
  const size_t n = (size_t)gridDim.y * L * L * H;
  (void)n;
  const int bl = blockIdx.y;
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < (size_t)L * L * H; o += (size_t)gridDim.x * blockDim.x) {
    const int h = o % H;
    const size_t ij = o / H;
    const int j = ij % L, i = ij / L;
    out[(size_t)(b0 + bl) * L * L * H + o] = alpha[((size_t)(bl * H + h) * L + i) * Lp + j];
  }
}

// ------------------------------------------------------------------------------------------ launchers
cudaError_t attn_kernels_init() { return cudaSuccess; }

void launch_alpha_tap(int nb, int b0, int L, int Lp, const float* alpha, float* out, cudaStream_t st) {
  ProfScope prof__(KK_OTHER, st);
  dim3 grid(64, nb);
  alpha_to_reference_layout<<<grid, 256, 0, st>>>(L, Lp, b0, alpha, out);
}

}  // namespace abopt
