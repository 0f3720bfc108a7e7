// Shared device/host helpers for libabopt_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libabopt_b200 is written for sm_100a (B200) only"
#endif

namespace abopt {

constexpr int F = 128;          // node channels
constexpr int C = 64;           // pair channels
constexpr int H = 12;           // heads
constexpr int D = 32;           // qk / value channels per head
constexpr int P = 8;            // points per head
constexpr int NPROJ = 3 * H * D + 3 * H * P * 3;   // 2016 projection columns: q | k | v | qp | kp | vp
constexpr int OFF_Q = 0, OFF_K = H * D, OFF_V = 2 * H * D;
constexpr int OFF_QP = 3 * H * D, OFF_KP = OFF_QP + H * P * 3, OFF_VP = OFF_KP + H * P * 3;
constexpr int NFEAT = H * C + H * D + H * P * 7;   // 1824 aggregate columns
constexpr int FEAT_NODE = H * C;                   // 768
constexpr int FEAT_PTS = FEAT_NODE + H * D;        // 1152
constexpr int FEAT_DIST = FEAT_PTS + H * P * 3;    // 1440
constexpr int FEAT_DIR = FEAT_DIST + H * P;        // 1536
constexpr int NAA = 20;
constexpr int NBINS = 8192;

// launch bookkeeping (abopt_kernel_launch_count) and the optional per-kernel event profiler
extern unsigned long long g_launches;
enum KernelKind { KK_MIXER = 0, KK_PROJ, KK_LOGITS, KK_PAIR, KK_AGGR, KK_TAIL, KK_HEADS, KK_STEP, KK_OTHER, KK_CTX, KK_PAIR_PART, KK_COUNT };
void prof_begin(int kind, cudaStream_t st);     // no-ops unless abopt_profile_enable(1)
void prof_end(int kind, cudaStream_t st);
struct ProfScope {
  int kind; cudaStream_t st;
  ProfScope(int k, cudaStream_t s) : kind(k), st(s) { prof_begin(k, s); }
  ~ProfScope() { prof_end(kind, st); g_launches += 1ull; }
};
inline void count_launch(int n = 1) { g_launches += (unsigned long long)n; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// x - trunc_tf32(x): the "lo" plane of the 3xTF32 operand split (the tensor core reads trunc_tf32(x) from the raw value)
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// ---------------------------------------------------------------- dihedral_from_four_points (modules/common/geometry.py:254-271)
// Signed angle between the planes (p0,p1,p2) and (p1,p2,p3), operation by operation as the reference evaluates it: products and
// sums are rounded separately (no FMA contraction) because the sign of the result is the sign of a triple product, which can be
// zero in exact arithmetic; cosine clamped to +-0.999999; NaN (degenerate normals: 0 / 0) -> 0 like torch.nan_to_num.
__device__ __forceinline__ void cross3_rn(const float* a, const float* b, float* o) {
  o[0] = __fsub_rn(__fmul_rn(a[1], b[2]), __fmul_rn(a[2], b[1]));
  o[1] = __fsub_rn(__fmul_rn(a[2], b[0]), __fmul_rn(a[0], b[2]));
  o[2] = __fsub_rn(__fmul_rn(a[0], b[1]), __fmul_rn(a[1], b[0]));
}
__device__ __forceinline__ float dot3_rn(const float* a, const float* b) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a[0], b[0]), __fmul_rn(a[1], b[1])), __fmul_rn(a[2], b[2]));
}
__device__ __forceinline__ float dihedral4(const float* p0, const float* p1, const float* p2, const float* p3) {
  float v0[3], v1[3], v2[3], u1[3], u2[3], w[3], n1[3], n2[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { v0[c] = p2[c] - p1[c]; v1[c] = p0[c] - p1[c]; v2[c] = p3[c] - p2[c]; }
  cross3_rn(v0, v1, u1);
  cross3_rn(v0, v2, u2);
  const float l1 = sqrtf(dot3_rn(u1, u1)), l2 = sqrtf(dot3_rn(u2, u2));
#pragma unroll
  for (int c = 0; c < 3; ++c) { n1[c] = u1[c] / l1; n2[c] = u2[c] / l2; }      // 0 / 0 -> NaN, as in the reference
  cross3_rn(v1, v2, w);
  const float tp = dot3_rn(w, v0);
  const float sgn = tp > 0.f ? 1.f : (tp < 0.f ? -1.f : 0.f);
  float cs = dot3_rn(n1, n2);
  if (isnan(cs) || isnan(tp)) return 0.f;                                       // nan_to_num, geometry.py:270
  cs = fminf(fmaxf(cs, -0.999999f), 0.999999f);
  return sgn * acosf(cs);
}

// ---------------------------------------------------------------- 3x3 helpers (row-major R[9])
struct Mat3 { float m[9]; };

__device__ __forceinline__ Mat3 matmul3(const Mat3& A, const Mat3& B) {
  Mat3 Cm;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      Cm.m[i * 3 + j] = A.m[i * 3 + 0] * B.m[0 * 3 + j] + A.m[i * 3 + 1] * B.m[1 * 3 + j] + A.m[i * 3 + 2] * B.m[2 * 3 + j];
  return Cm;
}

// Rodrigues exponential with the reference's guards (modules/common/so3.py:33-57).
__device__ __forceinline__ Mat3 so3_exp(float x, float y, float z) {
  // S = [[0, z, -y], [-z, 0, x], [y, -x, 0]]
  const float ang = sqrtf(x * x + y * y + z * z);
  const float b = (sinf(ang) + 1e-8f) / (ang + 1e-8f);
  const float c = (1.f - cosf(ang) + 1e-8f) / (ang * ang + 2e-8f);
  Mat3 S = {{0.f, z, -y, -z, 0.f, x, y, -x, 0.f}};
  Mat3 S2 = matmul3(S, S);
  Mat3 R;
#pragma unroll
  for (int i = 0; i < 9; ++i) R.m[i] = ((i % 4 == 0) ? 1.f : 0.f) + b * S.m[i] + c * S2.m[i];
  return R;
}

// Log map, op-for-op as the reference under no_grad (modules/common/so3.py:10-30): the
// sqrt(1 - cos^2) / acos formulation is kept on purpose (ill-conditioned near pi, SURVEY finding 4).
// `lo`: the cosine clamp, -1 under torch.no_grad() (sampling, validation) and -0.999 with autograd enabled (training), so3.py:12-17.
__device__ __forceinline__ void so3_log(const Mat3& R, float& x, float& y, float& z, float lo = -1.f) {
  const float tr = R.m[0] + R.m[4] + R.m[8];
  const float cos_t = fmaxf((tr - 1.f) / 2.f, lo);
  const float sin_t = sqrtf(1.f - cos_t * cos_t);
  const float theta = acosf(cos_t);
  const float coef = (theta + 1e-8f) / (2.f * sin_t + 2e-8f);
  x = coef * (R.m[1 * 3 + 2] - R.m[2 * 3 + 1]);
  y = coef * (R.m[2 * 3 + 0] - R.m[0 * 3 + 2]);
  z = coef * (R.m[0 * 3 + 1] - R.m[1 * 3 + 0]);
}

// (1 + b i + c j + d k) / |.| -> rotation (modules/common/geometry.py:215-233)
__device__ __forceinline__ Mat3 quat1ijk_to_rot(float b, float c, float d) {
  const float s = sqrtf(1.f + b * b + c * c + d * d);
  const float a = 1.f / s;
  b = b / s; c = c / s; d = d / s;
  Mat3 R = {{a * a + b * b - c * c - d * d, 2 * b * c - 2 * a * d, 2 * b * d + 2 * a * c,
             2 * b * c + 2 * a * d, a * a - b * b + c * c - d * d, 2 * c * d - 2 * a * b,
             2 * b * d - 2 * a * c, 2 * c * d + 2 * a * b, a * a - b * b - c * c + d * d}};
  return R;
}

// real-first quaternion, F.normalize'd twice as the reference does (geometry.py:148-174)
__device__ __forceinline__ Mat3 quat_to_rot(float r, float i, float j, float k) {
  float n = fmaxf(sqrtf(r * r + i * i + j * j + k * k), 1e-12f);
  r /= n; i /= n; j /= n; k /= n;
  n = fmaxf(sqrtf(r * r + i * i + j * j + k * k), 1e-12f);     // second F.normalize inside quaternion_to_rotation_matrix
  r /= n; i /= n; j /= n; k /= n;
  const float two_s = 2.0f / (r * r + i * i + j * j + k * k);
  Mat3 R = {{1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
             two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
             two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)}};
  return R;
}

// ---------------------------------------------------------------- Philox4x32-10 (fast-mode RNG)
struct Philox {
  uint32_t k0, k1;
  __device__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      const uint32_t n0 = hi1 ^ c1 ^ a, n2 = hi0 ^ c3 ^ b;
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
__device__ __forceinline__ float u01_open(uint32_t x) {      // (0, 1]
  return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ float u01_half(uint32_t x) {      // [0, 1)
  return (float)(x >> 8) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& g0, float& g1) {
  const float r = sqrtf(-2.0f * logf(u01_open(a)));
  float s, c;
  sincosf(6.28318530717958647692f * u01_half(b), &s, &c);
  g0 = r * c; g1 = r * s;
}

}  // namespace abopt
