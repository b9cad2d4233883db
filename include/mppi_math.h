/*
 * mppi_math.h — deterministic scalar math shared by the CUDA kernels and the CPU oracle.
 *
 * Why this exists (SURVEY.md §7 "Hard parts"): the integer costmap indices a
 * trajectory touches are a truncation of positions that come out of a chain of T sequential
 * float adds of vx*cos(yaw)*dt.  To compare indices bit-exactly between the GPU and the CPU
 * restatement, both sides must evaluate *the same* sin/cos and must not let the compiler fuse
 * mul+add differently.  Everything in this header is written with explicit fmaf() (single
 * rounding on both CPU and GPU) and plain IEEE +,-,*,/ so a build with contraction disabled
 * (`nvcc -fmad=false`, `g++ -ffp-contract=off`) produces identical bits on both sides.
 *
 * The reference gets sin/cos from Eigen's vectorised psin/pcos (optimizer.cpp:552-553), which is
 * itself a ~1-2 ulp polynomial whose bits depend on the Eigen version and ISA; it cannot be
 * reproduced here (Eigen is not in the container).  This routine has the same accuracy class.
 *
 * No ROS/Eigen/torch types; usable from C++ and CUDA.
 */
#ifndef MPPI_MATH_H_
#define MPPI_MATH_H_

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MPPI_HD __host__ __device__ __forceinline__
#else
#define MPPI_HD inline
#endif

#define MPPI_PIF 3.141592653589793238462643383279502884e+00F /* utils.hpp:50 M_PIF */
#define MPPI_PIF_2 1.5707963267948966e+00F                  /* utils.hpp:51 M_PIF_2 */

/* Cody-Waite split of pi/2 (three float terms). */
#define MPPI_PIO2_HI 1.5707962513e+00f
#define MPPI_PIO2_MID 7.5497894159e-08f
#define MPPI_PIO2_LO 5.3903029534e-15f
#define MPPI_2_OVER_PI 6.3661977237e-01f

/* The arithmetic of mppi_sincosf for an argument already known to satisfy |x| < 1e8 (the caller checks; the
 * rollout kernel tests the heading once per step and takes mppi_sincosf itself in the rare other case). */
MPPI_HD void mppi_sincosf_in_range(float xs, float * s_out, float * c_out)
{
  const float k = rintf(xs * MPPI_2_OVER_PI);
  float r = fmaf(-k, MPPI_PIO2_HI, xs);
  r = fmaf(-k, MPPI_PIO2_MID, r);
  r = fmaf(-k, MPPI_PIO2_LO, r);
  const int q = static_cast<int>(k) & 3;
  const float z = r * r;
  float ps = fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
  ps = fmaf(ps, z, -1.6666654611e-1f);
  const float sn = fmaf(ps * z, r, r);
  float pc = fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
  pc = fmaf(pc, z, 4.166664568298827e-2f);
  const float cs = fmaf(pc * z, z, fmaf(-0.5f, z, 1.0f));
  const float a = (q & 1) ? cs : sn;
  const float b = (q & 1) ? sn : cs;
  *s_out = (q & 2) ? -a : a;
  *c_out = ((q + 1) & 2) ? -b : b;
}

/* sin and cos of x, ~1 ulp for |x| < ~1e4, deterministic everywhere (NaN for non-finite). */
MPPI_HD void mppi_sincosf(float x, float * s_out, float * c_out)
{
  /* huge, inf or nan: no meaningful heading; propagate NaN.  Selected at the end instead of an early
   * return so the routine is straight-line code (the GPU interleaves several calls for latency). */
  const bool ok = fabsf(x) < 1.0e8f;
  const float xs = ok ? x : 0.0f;
  const float k = rintf(xs * MPPI_2_OVER_PI);
  float r = fmaf(-k, MPPI_PIO2_HI, xs);
  r = fmaf(-k, MPPI_PIO2_MID, r);
  r = fmaf(-k, MPPI_PIO2_LO, r);
  const int q = static_cast<int>(k) & 3;
  const float z = r * r;
  /* minimax polynomials on [-pi/4, pi/4] (Cephes single-precision coefficients) */
  float ps = fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
  ps = fmaf(ps, z, -1.6666654611e-1f);
  const float sn = fmaf(ps * z, r, r);
  float pc = fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
  pc = fmaf(pc, z, 4.166664568298827e-2f);
  const float cs = fmaf(pc * z, z, fmaf(-0.5f, z, 1.0f));
  /* quadrant: q odd swaps sin/cos; sin is negated for q = 2, 3 and cos for q = 1, 2 */
  const float a = (q & 1) ? cs : sn;
  const float b = (q & 1) ? sn : cs;
  const float s = (q & 2) ? -a : a;
  const float c = ((q + 1) & 2) ? -b : b;
  const float bad = x - x;  /* 0 for finite, NaN otherwise */
  *s_out = ok ? s : bad;
  *c_out = ok ? c : 1.0f + bad;
}

/* utils.hpp:254-262 normalize_angles (float, Eigen side): fmodf is exact on CPU and GPU. */
/* fmodf(x, m) for m > 0 without the library's division loop when |x| < 2m, which is every angle this code meets:
 * fmod is exact and keeps the dividend's sign, so it is x itself for |x| < m and x -/+ m for m <= |x| < 2m — and that
 * difference is exactly representable (Sterbenz: m <= |x| <= 2m), so the float subtraction returns the same bits. */
MPPI_HD float mppi_fmodf_small(float x, float m)
{
  const float ax = fabsf(x);
  if (ax < m) {return x;}
  if (ax < 2.0f * m) {
    const float r = x < 0.0f ? x + m : x - m;
    return r == 0.0f ? copysignf(0.0f, x) : r;   /* fmod(-m, m) is -0 */
  }
  return fmodf(x, m);
}
MPPI_HD double mppi_fmod_small(double x, double m)
{
  const double ax = fabs(x);
  if (ax < m) {return x;}
  if (ax < 2.0 * m) {
    const double r = x < 0.0 ? x + m : x - m;
    return r == 0.0 ? copysign(0.0, x) : r;
  }
  return fmod(x, m);
}

MPPI_HD float mppi_normalize_anglef(float angle)
{
  const float x = angle + MPPI_PIF;
  const float rem = mppi_fmodf_small(x, 2.0f * MPPI_PIF);
  return rem < 0.0f ? rem + MPPI_PIF : rem - MPPI_PIF;
}

/* utils.hpp:277-283 shortest_angular_distance(from, to) = normalize_angles(to - from). */
MPPI_HD float mppi_shortest_angular_distancef(float from, float to)
{
  return mppi_normalize_anglef(to - from);
}

/*
 * ros/angles (branch ros2, not vendored under /root/reference; pinned only as a comment in
 * tools/underlay.repos) angles::normalize_angle(double), restated from the published header:
 *   result = fmod(angle + pi, 2pi); if (result <= 0) return result + pi; return result - pi;
 */
MPPI_HD double mppi_angles_normalize_angle(double angle)
{
  const double result = mppi_fmod_small(angle + M_PI, 2.0 * M_PI);
  if (result <= 0.0) {return result + M_PI;}
  return result - M_PI;
}

/* ros/angles angles::shortest_angular_distance(from, to) = normalize_angle(to - from). */
MPPI_HD double mppi_angles_shortest_angular_distance(double from, double to)
{
  return mppi_angles_normalize_angle(to - from);
}

/* x^p for unsigned integer p by repeated multiplication (Eigen integer-exponent pow). */
MPPI_HD float mppi_powu(float x, unsigned int p)
{
  float result = 1.0f;
  float base = x;
  while (p) {
    if (p & 1u) {result *= base;}
    p >>= 1u;
    if (p) {base *= base;}
  }
  return result;
}

/* utils.hpp:659-663 clamp(lower, upper, input) = min(upper, max(input, lower)). */
MPPI_HD float mppi_clampf(float lower_bound, float upper_bound, float input)
{
  return fminf(upper_bound, fmaxf(input, lower_bound));
}

/* utils.hpp:674-682 clampVelocityByAccel: note `>= 0` (predict() uses `> 0`). */
MPPI_HD float mppi_clamp_velocity_by_accel(
  float last_vel, float curr_vel, float min_delta, float max_delta)
{
  if (last_vel >= 0.0f) {
    return mppi_clampf(last_vel + min_delta, last_vel + max_delta, curr_vel);
  }
  return mppi_clampf(last_vel - max_delta, last_vel - min_delta, curr_vel);
}

#endif  /* MPPI_MATH_H_ */
