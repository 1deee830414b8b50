// Minimal stand-in for the CUDA-samples "helper_math.h" that the reference sources include
// (the samples header is not part of the reference repository nor of this image; see
// /root/reference/README.md:17).  TEST INFRASTRUCTURE ONLY: used to compile the unmodified
// reference .cu files into oracle/_ref/.  Only the overloads the reference actually uses are
// provided; each one evaluates component-wise in the obvious left-to-right order so that the
// float rounding sequence is the one the samples header defines.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define HM_FN inline __host__ __device__

// ---- constructors -------------------------------------------------------------------------
HM_FN float3 make_float3(float s) { return make_float3(s, s, s); }
HM_FN float3 make_float3(int3 v) { return make_float3(float(v.x), float(v.y), float(v.z)); }
HM_FN int3 make_int3(int s) { return make_int3(s, s, s); }
HM_FN int3 make_int3(float3 v) { return make_int3(int(v.x), int(v.y), int(v.z)); }

// ---- float3 arithmetic --------------------------------------------------------------------
HM_FN float3 operator-(float3 v) { return make_float3(-v.x, -v.y, -v.z); }
HM_FN float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
HM_FN float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
HM_FN float3 operator*(float3 a, float3 b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
HM_FN float3 operator/(float3 a, float3 b) { return make_float3(a.x / b.x, a.y / b.y, a.z / b.z); }
HM_FN float3 operator+(float3 a, float s) { return make_float3(a.x + s, a.y + s, a.z + s); }
HM_FN float3 operator-(float3 a, float s) { return make_float3(a.x - s, a.y - s, a.z - s); }
HM_FN float3 operator*(float3 a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
HM_FN float3 operator*(float s, float3 a) { return make_float3(s * a.x, s * a.y, s * a.z); }
HM_FN float3 operator/(float3 a, float s) { return make_float3(a.x / s, a.y / s, a.z / s); }
HM_FN float3 operator/(float s, float3 a) { return make_float3(s / a.x, s / a.y, s / a.z); }
HM_FN void operator+=(float3& a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
HM_FN void operator-=(float3& a, float3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
HM_FN void operator*=(float3& a, float s) { a.x *= s; a.y *= s; a.z *= s; }

// ---- int3 arithmetic ----------------------------------------------------------------------
HM_FN int3 operator+(int3 a, int3 b) { return make_int3(a.x + b.x, a.y + b.y, a.z + b.z); }
HM_FN int3 operator-(int3 a, int3 b) { return make_int3(a.x - b.x, a.y - b.y, a.z - b.z); }
HM_FN int3 operator*(int s, int3 a) { return make_int3(s * a.x, s * a.y, s * a.z); }
HM_FN int3 operator*(int3 a, int s) { return make_int3(a.x * s, a.y * s, a.z * s); }

// ---- geometry -----------------------------------------------------------------------------
#ifndef __CUDACC__
// host-only parse of the reference HEADERS by g++ (tests/api_conformance.cpp); never executed
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
#endif
HM_FN float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
HM_FN float length(float3 v) { return sqrtf(dot(v, v)); }
HM_FN float3 normalize(float3 v) { return v * rsqrtf(dot(v, v)); }

#ifndef __CUDACC__
inline int max(int a, int b) { return a > b ? a : b; }
#endif
