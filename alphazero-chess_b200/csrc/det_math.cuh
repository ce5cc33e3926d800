// IEEE-exact log / exp built only from + - x / (no FMA contraction: the including translation units are compiled with
// -fmad=false), bit-identical to oracle/mcts.cpp (spec in DESIGN.md section 3), and the improved-policy power of tree.rs:174.
#pragma once
#include "chess.cuh"

namespace azb {

__device__ inline double det_log(double x) {
    u64 bits = (u64)__double_as_longlong(x);
    int e = (int)((bits >> 52) & 0x7FF) - 1023;
    bits = (bits & 0x000FFFFFFFFFFFFFULL) | 0x3FF0000000000000ULL;
    double m = __longlong_as_double((long long)bits);
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double s = (m - 1.0) / (m + 1.0);
    double s2 = s * s;
    double t = 1.0 / 23.0;
    t = t * s2 + 1.0 / 21.0; t = t * s2 + 1.0 / 19.0; t = t * s2 + 1.0 / 17.0; t = t * s2 + 1.0 / 15.0;
    t = t * s2 + 1.0 / 13.0; t = t * s2 + 1.0 / 11.0; t = t * s2 + 1.0 / 9.0; t = t * s2 + 1.0 / 7.0;
    t = t * s2 + 1.0 / 5.0; t = t * s2 + 1.0 / 3.0; t = t * s2 + 1.0;
    return (double)e * 0.6931471805599453 + 2.0 * s * t;
}
__device__ inline double det_exp(double x) {
    double kf = x * 1.4426950408889634;
    long long k = (long long)(kf < 0 ? kf - 0.5 : kf + 0.5);
    double r = x - (double)k * 0.6931471805599453;
    double t = 1.0 / 6227020800.0;
    t = t * r + 1.0 / 479001600.0; t = t * r + 1.0 / 39916800.0; t = t * r + 1.0 / 3628800.0;
    t = t * r + 1.0 / 362880.0; t = t * r + 1.0 / 40320.0; t = t * r + 1.0 / 5040.0; t = t * r + 1.0 / 720.0;
    t = t * r + 1.0 / 120.0; t = t * r + 1.0 / 24.0; t = t * r + 1.0 / 6.0; t = t * r + 0.5; t = t * r + 1.0;
    t = t * r + 1.0;
    double sc = __longlong_as_double((long long)((u64)(k + 1023) << 52));
    return t * sc;
}
// x^(1/T) for a visit count (tree.rs:174); identity at T = 1.  Restated with det_log / det_exp in f64 so that the oracle and
// the engine agree bit for bit (against libm's powf the last bit can differ when T != 1).
__device__ __forceinline__ float pow_inv_temperature(float visits, float inv_t) {
    if (inv_t == 1.0f || visits == 0.0f) return visits;
    return (float)det_exp(det_log((double)visits) * (double)inv_t);
}

}  // namespace azb
