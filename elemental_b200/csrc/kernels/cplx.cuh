// Minimal POD complex type usable in device code, layout-compatible with
// std::complex<R> / El::Complex<R> (interleaved re,im).
#pragma once
#include <cuda_runtime.h>

namespace elb200 {

template <class R>
struct __align__(sizeof(R) * 2) cplx {
    R re, im;
};

template <class R> __host__ __device__ inline cplx<R> mk(R a, R b) { cplx<R> z; z.re = a; z.im = b; return z; }
template <class R> __host__ __device__ inline cplx<R> operator+(cplx<R> a, cplx<R> b) { return mk(a.re + b.re, a.im + b.im); }
template <class R> __host__ __device__ inline cplx<R> operator-(cplx<R> a, cplx<R> b) { return mk(a.re - b.re, a.im - b.im); }
template <class R> __host__ __device__ inline cplx<R> operator-(cplx<R> a) { return mk(-a.re, -a.im); }
template <class R> __host__ __device__ inline cplx<R> operator*(cplx<R> a, cplx<R> b) {
    return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
template <class R> __host__ __device__ inline cplx<R> operator*(R a, cplx<R> b) { return mk(a * b.re, a * b.im); }
template <class R> __host__ __device__ inline cplx<R> operator*(cplx<R> a, R b) { return mk(a.re * b, a.im * b); }
template <class R> __host__ __device__ inline cplx<R>& operator+=(cplx<R>& a, cplx<R> b) { a.re += b.re; a.im += b.im; return a; }
template <class R> __host__ __device__ inline cplx<R>& operator-=(cplx<R>& a, cplx<R> b) { a.re -= b.re; a.im -= b.im; return a; }

// Scalar traits shared by real and complex kernels
template <class T> struct scalar_traits;
template <> struct scalar_traits<float> {
    typedef float real;
    static constexpr bool is_complex = false;
    __host__ __device__ static float conj(float x) { return x; }
    __host__ __device__ static float zero() { return 0.f; }
    __host__ __device__ static float from_real(float x) { return x; }
    __host__ __device__ static float real_part(float x) { return x; }
    __host__ __device__ static float abs2(float x) { return x * x; }
    __host__ __device__ static bool is_zero(float x) { return x == 0.f; }
    __host__ __device__ static bool is_one(float x) { return x == 1.f; }
};
template <> struct scalar_traits<double> {
    typedef double real;
    static constexpr bool is_complex = false;
    __host__ __device__ static double conj(double x) { return x; }
    __host__ __device__ static double zero() { return 0.0; }
    __host__ __device__ static double from_real(double x) { return x; }
    __host__ __device__ static double real_part(double x) { return x; }
    __host__ __device__ static double abs2(double x) { return x * x; }
    __host__ __device__ static bool is_zero(double x) { return x == 0.0; }
    __host__ __device__ static bool is_one(double x) { return x == 1.0; }
};
template <class R> struct scalar_traits<cplx<R>> {
    typedef R real;
    static constexpr bool is_complex = true;
    __host__ __device__ static cplx<R> conj(cplx<R> x) { return mk(x.re, -x.im); }
    __host__ __device__ static cplx<R> zero() { return mk(R(0), R(0)); }
    __host__ __device__ static cplx<R> from_real(R x) { return mk(x, R(0)); }
    __host__ __device__ static R real_part(cplx<R> x) { return x.re; }
    __host__ __device__ static R abs2(cplx<R> x) { return x.re * x.re + x.im * x.im; }
    __host__ __device__ static bool is_zero(cplx<R> x) { return x.re == R(0) && x.im == R(0); }
    __host__ __device__ static bool is_one(cplx<R> x) { return x.re == R(1) && x.im == R(0); }
};

typedef cplx<float> c32_t;
typedef cplx<double> c64_t;

}  // namespace elb200
