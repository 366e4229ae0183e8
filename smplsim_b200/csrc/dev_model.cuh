// dev_model.cuh -- small device math shared by the kernels: 3-vectors, wxyz quaternions, spatial 6-vectors, symmetric 6x6
// (upper triangle), rigid-body inertia about a reference point, Philox4x32-10.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/smplsim.h"

// ----------------------------------------------------------------------------- small math
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(float* p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }

struct Q4 { float w, x, y, z; };
__device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {
  Q4 r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  return r;
}
__device__ __forceinline__ Q4 qnormalize(Q4 q) {
  float n = sqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  if (n < 1e-15f) { q.w = 1.f; q.x = q.y = q.z = 0.f; return q; }
  float s = 1.0f / n;
  q.w *= s; q.x *= s; q.y *= s; q.z *= s;
  return q;
}
// rotate v by unit quaternion q
__device__ __forceinline__ V3 qrot(Q4 q, V3 v) {
  V3 u = v3(q.x, q.y, q.z);
  V3 t = 2.0f * cross(u, v);
  return v + q.w * t + cross(u, t);
}
// the reference's quat_rotate (np_transform_utils.py:23-32): v(2w^2-1) + 2w(q x v) + 2q(q.v); valid for any unit q
__device__ __forceinline__ V3 qrot_ref(Q4 q, V3 v) {
  V3 u = v3(q.x, q.y, q.z);
  V3 c = cross(u, v);
  float d = dot(u, v), s = 2.0f * q.w * q.w - 1.0f;
  return s * v + (2.0f * q.w) * c + (2.0f * d) * u;
}
__device__ __forceinline__ void q2mat(Q4 q, float* m) {
  float w = q.w, x = q.x, y = q.y, z = q.z;
  m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = 1 - 2 * (x * x + z * z); m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = 1 - 2 * (x * x + y * y);
}
__device__ __forceinline__ V3 mrot(const float* m, V3 v) {
  return v3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z);
}

// spatial 6-vectors: motion [w; v], force [n; f]
struct S6 { V3 a, l; };  // angular / linear part
__device__ __forceinline__ S6 s6(V3 a, V3 l) { S6 r; r.a = a; r.l = l; return r; }
__device__ __forceinline__ S6 operator+(S6 p, S6 q) { return s6(p.a + q.a, p.l + q.l); }
__device__ __forceinline__ S6 operator-(S6 p, S6 q) { return s6(p.a - q.a, p.l - q.l); }
__device__ __forceinline__ S6 operator*(float s, S6 p) { return s6(s * p.a, s * p.l); }
__device__ __forceinline__ float dot6(S6 p, S6 q) { return dot(p.a, q.a) + dot(p.l, q.l); }
__device__ __forceinline__ S6 ld6(const float* p) { return s6(ld3(p), ld3(p + 3)); }
__device__ __forceinline__ void st6(float* p, S6 a) { st3(p, a.a); st3(p + 3, a.l); }
__device__ __forceinline__ S6 cross_motion(S6 v, S6 s) { return s6(cross(v.a, s.a), cross(v.a, s.l) + cross(v.l, s.a)); }
__device__ __forceinline__ S6 cross_force(S6 v, S6 f) { return s6(cross(v.a, f.a) + cross(v.l, f.l), cross(v.a, f.l)); }

// symmetric 6x6, upper triangle row-major: (i,j), i<=j -> i*6 - i*(i-1)/2 + (j-i)
__host__ __device__ constexpr int sidx(int i, int j) { return i <= j ? (i * 6 - (i * (i - 1)) / 2 + (j - i)) : (j * 6 - (j * (j - 1)) / 2 + (i - j)); }

__device__ __forceinline__ void sym_mul(const float* A, const float* s, float* y) {
#pragma unroll
  for (int i = 0; i < 6; i++) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 6; j++) acc = fmaf(A[sidx(i, j)], s[j], acc);
    y[i] = acc;
  }
}
// A -= c * u u^T
__device__ __forceinline__ void sym_rank1(float* A, const float* u, float c) {
#pragma unroll
  for (int i = 0; i < 6; i++) {
    float ci = c * u[i];
#pragma unroll
    for (int j = i; j < 6; j++) A[sidx(i, j)] = fmaf(-ci, u[j], A[sidx(i, j)]);
  }
}
// rigid-body inertia about the reference point: m, c = m r, I (xx yy zz xy xz yz)
__device__ __forceinline__ void rb_expand(const float* r10, float* A) {
  float m = r10[0], cx = r10[1], cy = r10[2], cz = r10[3];
  A[sidx(0, 0)] = r10[4]; A[sidx(0, 1)] = r10[7]; A[sidx(0, 2)] = r10[8]; A[sidx(0, 3)] = 0.f; A[sidx(0, 4)] = -cz; A[sidx(0, 5)] = cy;
  A[sidx(1, 1)] = r10[5]; A[sidx(1, 2)] = r10[9]; A[sidx(1, 3)] = cz; A[sidx(1, 4)] = 0.f; A[sidx(1, 5)] = -cx;
  A[sidx(2, 2)] = r10[6]; A[sidx(2, 3)] = -cy; A[sidx(2, 4)] = cx; A[sidx(2, 5)] = 0.f;
  A[sidx(3, 3)] = m; A[sidx(3, 4)] = 0.f; A[sidx(3, 5)] = 0.f; A[sidx(4, 4)] = m; A[sidx(4, 5)] = 0.f; A[sidx(5, 5)] = m;
}
__device__ __forceinline__ S6 rb_mul(const float* r10, S6 v) {
  float m = r10[0];
  V3 c = v3(r10[1], r10[2], r10[3]);
  V3 Iw = v3(r10[4] * v.a.x + r10[7] * v.a.y + r10[8] * v.a.z, r10[7] * v.a.x + r10[5] * v.a.y + r10[9] * v.a.z,
             r10[8] * v.a.x + r10[9] * v.a.y + r10[6] * v.a.z);
  return s6(Iw + cross(c, v.l), m * v.l - cross(c, v.a));
}

// Philox4x32-10 (same stream as oracle/mjstep_oracle.c)
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    uint32_t n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
__device__ __forceinline__ int rand_range(uint32_t x, int lo, int hi) { return lo + (int)__umulhi(x, (uint32_t)(hi - lo)); }
