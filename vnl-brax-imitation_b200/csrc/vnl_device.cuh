// vnl_device.cuh -- small fp32 device helpers (quaternions, spatial algebra, reductions).
// Spatial vectors are [angular(3), linear(3)]; quaternions are [w, x, y, z]; cinert is
// [Ixx Iyy Izz Ixy Ixz Iyz, m*cx m*cy m*cz, m] (MJX `smooth.com_pos` layout).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define VNL_MINVAL 1e-15f
#define VNL_MINIMP 1e-4f
#define VNL_MAXIMP 0.9999f

struct V3 { float x, y, z; };
struct Q4 { float w, x, y, z; };

__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(float* p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
__device__ __forceinline__ Q4 ld4(const float* p) { Q4 r; r.w = p[0]; r.x = p[1]; r.y = p[2]; r.z = p[3]; return r; }
__device__ __forceinline__ void st4(float* p, Q4 q) { p[0] = q.w; p[1] = q.x; p[2] = q.y; p[3] = q.z; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

__device__ __forceinline__ Q4 quat_mul(Q4 u, Q4 v) {
  Q4 r;
  r.w = u.w * v.w - u.x * v.x - u.y * v.y - u.z * v.z;
  r.x = u.w * v.x + u.x * v.w + u.y * v.z - u.z * v.y;
  r.y = u.w * v.y - u.x * v.z + u.y * v.w + u.z * v.x;
  r.z = u.w * v.z + u.x * v.y - u.y * v.x + u.z * v.w;
  return r;
}
// MJX math.rotate: 2 (u.v) u + (s^2 - u.u) v + 2 s (u x v)
__device__ __forceinline__ V3 rotate(V3 vec, Q4 q) {
  V3 u = v3(q.x, q.y, q.z);
  float s = q.w, ud = dot(u, vec), uu = dot(u, u);
  V3 c = cross(u, vec);
  return u * (2.0f * ud) + vec * (s * s - uu) + c * (2.0f * s);
}
__device__ __forceinline__ void quat_to_mat(Q4 q, float* m) {  // row-major 3x3
  float q00 = q.w * q.w, q01 = q.w * q.x, q02 = q.w * q.y, q03 = q.w * q.z;
  float q11 = q.x * q.x, q12 = q.x * q.y, q13 = q.x * q.z, q22 = q.y * q.y, q23 = q.y * q.z, q33 = q.z * q.z;
  m[0] = q00 + q11 - q22 - q33; m[1] = 2.0f * (q12 - q03); m[2] = 2.0f * (q13 + q02);
  m[3] = 2.0f * (q12 + q03); m[4] = q00 - q11 + q22 - q33; m[5] = 2.0f * (q23 - q01);
  m[6] = 2.0f * (q13 - q02); m[7] = 2.0f * (q23 + q01); m[8] = q00 - q11 - q22 + q33;
}
__device__ __forceinline__ Q4 axis_angle_quat(V3 axis, float angle) {
  float s, c;
  sincosf(angle * 0.5f, &s, &c);
  Q4 q; q.w = c; q.x = axis.x * s; q.y = axis.y * s; q.z = axis.z * s;
  return q;
}
__device__ __forceinline__ Q4 quat_normalize(Q4 q) {
  float n = sqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  if (n == 0.0f) return q;
  q.w = q.w / n; q.x = q.x / n; q.y = q.y / n; q.z = q.z / n;
  return q;
}
__device__ __forceinline__ float normalize3(V3& v) {
  float n = sqrtf(dot(v, v));
  if (n == 0.0f) return n;
  v.x /= n; v.y /= n; v.z /= n;
  return n;
}

// r = I (x) v for a cinert-format spatial inertia
__device__ __forceinline__ void inert_mul(const float* i, const float* v, float* r) {
  V3 va = ld3(v), vl = ld3(v + 3), mc = ld3(i + 6);
  V3 ang = v3(i[0] * va.x + i[3] * va.y + i[4] * va.z, i[3] * va.x + i[1] * va.y + i[5] * va.z, i[4] * va.x + i[5] * va.y + i[2] * va.z);
  st3(r, ang + cross(mc, vl));
  st3(r + 3, vl * i[9] - cross(mc, va));
}
__device__ __forceinline__ void motion_cross(const float* u, const float* v, float* r) {
  V3 ua = ld3(u), ul = ld3(u + 3), va = ld3(v), vl = ld3(v + 3);
  st3(r, cross(ua, va));
  st3(r + 3, cross(ul, va) + cross(ua, vl));
}
__device__ __forceinline__ void motion_cross_force(const float* v, const float* f, float* r) {
  V3 va = ld3(v), vl = ld3(v + 3), fa = ld3(f), fl = ld3(f + 3);
  st3(r, cross(va, fa) + cross(vl, fl));
  st3(r + 3, cross(va, fl));
}
__device__ __forceinline__ float dot6(const float* a, const float* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// N = 2, 4 or 8 sums at once, every lane ends with all N totals -- BIT-IDENTICAL to N calls of warp_sum (same pairing order
// 16, 8, 4, 2, 1; fp addition commutes), with 2 N + 4 - log2 N shuffles instead of 5 N: a reduce-scatter (each xor step halves the
// values a lane carries: it keeps one half and sends the other to its partner) down to one value per lane, the remaining
// butterfly steps on that value, and one indexed shuffle per total to hand them round.  8 sums: 17 shuffles instead of 40, 4 sums:
// 10 instead of 20.  The shuffle unit issues one warp instruction per cycle per SM and the lockstep puts all 14 warps of the SM in
// the same reduction at the same time, so the line search's 9 sums per evaluation were shuffle-throughput bound.
template <int N>
__device__ __forceinline__ void warp_sum_n(float (&v)[N]) {
  static_assert(N == 2 || N == 4 || N == 8, "warp_sum_n: N = 2, 4, 8");
  constexpr int LOG = N == 8 ? 3 : (N == 4 ? 2 : 1);
  const int lane = threadIdx.x & 31;
  int o = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1, o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int k = 0; k < n / 2; ++k) {
      const float send = up ? v[k] : v[k + n / 2];
      const float keep = up ? v[k + n / 2] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  float t = v[0];  // total-in-progress of value number (lane >> (5 - LOG))
#pragma unroll
  for (int q = 16 >> LOG; q > 0; q >>= 1) t += __shfl_xor_sync(0xffffffffu, t, q);
#pragma unroll
  for (int j = 0; j < N; ++j) v[j] = __shfl_sync(0xffffffffu, t, j << (5 - LOG));
}
