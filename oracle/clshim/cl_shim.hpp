// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// A C++ prelude under which the reference's OpenCL C kernel sources (kernel_farfield.cl, kernel_nearfield.cl,
// rendered for one dtype exactly as calc.py:605-624 renders them with Mako) compile with g++ and run one work-item
// at a time on the host.  It supplies only what those two files use: the address-space qualifiers, `uint`,
// the 3-vector types with component-wise arithmetic and `.s0/.s1/.s2`, `dot`, `rsqrt`, `sin/cos/sqrt/fabs` in the
// compute type, the `native_` spellings, and `get_global_id(0)`.
//
// Arithmetic policy -- the same one oracle_kernels.cpp fixes (OpenCL leaves these implementation-defined):
//   dot(a,b) = ((a0*b0 + a1*b1) + a2*b2);  rsqrt(x) = 1/sqrt(x);  sin/cos/sqrt from libm in the compute type;
//   contraction is decided by the compiler flag (-ffp-contract=off for the parity build).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>


typedef unsigned int uint;
#define __kernel static
#define __global
#define __constant const
#define __local

namespace clshim {
template <typename T> struct vec3 { T s0, s1, s2; };
template <typename T> inline vec3<T> operator+(vec3<T> p, vec3<T> q) { return {p.s0 + q.s0, p.s1 + q.s1, p.s2 + q.s2}; }
template <typename T> inline vec3<T> operator-(vec3<T> p, vec3<T> q) { return {p.s0 - q.s0, p.s1 - q.s1, p.s2 - q.s2}; }
template <typename T> inline vec3<T> operator-(vec3<T> p) { return {-p.s0, -p.s1, -p.s2}; }
template <typename T> inline vec3<T> operator*(vec3<T> p, vec3<T> q) { return {p.s0 * q.s0, p.s1 * q.s1, p.s2 * q.s2}; }
template <typename T> inline vec3<T> operator/(vec3<T> p, vec3<T> q) { return {p.s0 / q.s0, p.s1 / q.s1, p.s2 / q.s2}; }
// OpenCL widens the scalar operand to the vector's element type
template <typename T> inline vec3<T> operator*(T s, vec3<T> v) { return {s * v.s0, s * v.s1, s * v.s2}; }
template <typename T> inline vec3<T> operator*(vec3<T> v, T s) { return {v.s0 * s, v.s1 * s, v.s2 * s}; }
template <typename T> inline vec3<T> operator/(vec3<T> v, T s) { return {v.s0 / s, v.s1 / s, v.s2 / s}; }
template <typename T> inline vec3<T> operator+(vec3<T> v, T s) { return {v.s0 + s, v.s1 + s, v.s2 + s}; }
template <typename T> inline vec3<T> operator-(vec3<T> v, T s) { return {v.s0 - s, v.s1 - s, v.s2 - s}; }
template <typename T> inline vec3<T>& operator+=(vec3<T>& p, vec3<T> q) { p = p + q; return p; }
template <typename T> inline vec3<T>& operator-=(vec3<T>& p, vec3<T> q) { p = p - q; return p; }
template <typename T> inline vec3<T>& operator*=(vec3<T>& p, T s) { p = p * s; return p; }
template <typename T> inline vec3<T>& operator*=(vec3<T>& p, vec3<T> q) { p = p * q; return p; }

template <typename T> inline T cl_dot(vec3<T> p, vec3<T> q) { return (p.s0 * q.s0 + p.s1 * q.s1) + p.s2 * q.s2; }
inline double cl_sin(double x) { return ::sin(x); }
inline double cl_cos(double x) { return ::cos(x); }
inline double cl_sqrt(double x) { return ::sqrt(x); }
inline double cl_fabs(double x) { return ::fabs(x); }
inline float cl_sin(float x) { return ::sinf(x); }
inline float cl_cos(float x) { return ::cosf(x); }
inline float cl_sqrt(float x) { return ::sqrtf(x); }
inline float cl_fabs(float x) { return ::fabsf(x); }
template <typename T> inline T cl_rsqrt(T x) { return (T)1 / cl_sqrt(x); }

extern thread_local size_t global_id0;
inline size_t get_global_id(int) { return global_id0; }
}  // namespace clshim

typedef clshim::vec3<double> double3;
typedef clshim::vec3<float> float3;
using clshim::get_global_id;

#define dot clshim::cl_dot
#define sin clshim::cl_sin
#define cos clshim::cl_cos
#define sqrt clshim::cl_sqrt
#define fabs clshim::cl_fabs
#define rsqrt clshim::cl_rsqrt
#define native_sin clshim::cl_sin
#define native_cos clshim::cl_cos
#define native_sqrt clshim::cl_sqrt
#define native_rsqrt clshim::cl_rsqrt
