// Warp-cooperative resolution of the pinv branch (osc.py:52-55) without an eigen-decomposition, for the
// fix-up kernel (osc_tail.cuh).  EXPERIMENTAL: validated against numpy's pinv on the CPU (the routine is
// written in warp-uniform phases so that tests/host_fused can emulate the 32 lanes), not yet measured on a
// GPU; osc_tail_fixup only calls it when the environment asks for it (IRLOSC_FIXUP_COOP=1).
//
// numpy's pinv(A, rcond) keeps an eigenvalue iff it is > rcond * lambda_max.  With rigorous bounds
//   L <= lambda_max <= U      repeated squaring of A / tr A: tr(B^(2^m))^(1/2^m) is within K^(1/2^m) of it
// three outcomes are certain without eigenvectors of the whole matrix:
//   1  1 / tr(A^-1) > rcond U : nothing is cut, w = A^-1 g;
//   2  the Rayleigh quotient rho of the inverse-iteration vector x (so lambda_min <= rho) is <= rcond L, and
//      A' = A + tr(A) x x^T has 1 / tr(A'^-1) > rcond U.  Eigenvalues interlace (lambda_2(A) >= lambda_1(A')),
//      so exactly one eigenvalue is cut: w = A'^-1 (g - x (x . g));
//   0  anything else (inside the bounds, two or more small eigenvalues, a failed pivot): the caller runs the
//      Jacobi eigen-solver as before.
// Why this and not the in-thread version tried in round 1: the same arithmetic as a
// divergent one-lane path over local memory cost ~1 ms per step.  Here rows live in shared memory, the K
// column solves of A^-1 run on K lanes at once, and the whole thing is two small factorisations, ~6 matrix
// squarings of a <= 13 x 13 matrix and a few mat-vecs - against ~8 Jacobi sweeps of 4 barriers per round.
#pragma once
#include <cmath>
#include "irlosc_device.cuh"

namespace irlosc {
namespace fused {

template <int K>
struct CoopSmem {
    double A[K][K + 1];      // input, preserved
    double L[K][K + 1];      // factors (strict lower part, unit diagonal implied)
    double Y[K][K + 1];      // inverse
    double B[K][K + 1], C[K][K + 1];
    double dinv[K], g[K], x[K], xn[K], t[K], w[K];
    int fail;
};

// executors: `par(f)` runs f(lane) for the 32 lanes of a phase and separates phases
struct CoopHostEx {
    template <class F>
    void par(F f) { for (int lane = 0; lane < 32; ++lane) f(lane); }
};
#ifdef __CUDACC__
struct CoopDevEx {
    int lane;
    template <class F>
    __device__ __forceinline__ void par(F f) { __syncwarp(); f(lane); __syncwarp(); }
};
#endif

// L D L^T of S.B (lower part) into S.L / S.dinv; false when a pivot is not positive.
template <int K, class EX>
IRLOSC_HD bool coop_factor(CoopSmem<K> &S, EX &ex) {
    ex.par([&](int lane) {
        for (int e = lane; e < K * K; e += 32) S.L[e / K][e % K] = S.B[e / K][e % K];
        if (lane == 0) S.fail = 0;
    });
    for (int p = 0; p < K; ++p) {
        const double d = S.L[p][p];
        if (!(d > 0.0)) return false;
        const double inv = 1.0 / d;
        ex.par([&](int lane) {
            const int i = lane;
            if (i > p && i < K) {
                const double l = S.L[i][p] * inv;
                for (int j = p + 1; j <= i; ++j) S.L[i][j] = fma(-l, S.L[j][p], S.L[i][j]);
            }
            if (lane == 0) S.dinv[p] = inv;
        });
        ex.par([&](int lane) {
            if (lane > p && lane < K) S.L[lane][p] *= inv;
        });
    }
    return true;
}

// S.Y = (L D L^T)^-1, one column per lane (the inverse is symmetric: row j of Y holds column j).
template <int K, class EX>
IRLOSC_HD void coop_inverse(CoopSmem<K> &S, EX &ex) {
    ex.par([&](int lane) {
        const int c = lane;
        if (c >= K) return;
        double *y = S.Y[c];
        for (int i = 0; i < K; ++i) {
            double z = (i == c) ? 1.0 : 0.0;
            for (int j = 0; j < i; ++j) z = fma(-S.L[i][j], y[j], z);
            y[i] = z;
        }
        for (int i = 0; i < K; ++i) y[i] *= S.dinv[i];
        for (int i = K - 1; i >= 0; --i) {
            double z = y[i];
            for (int j = i + 1; j < K; ++j) z = fma(-S.L[j][i], y[j], z);
            y[i] = z;
        }
    });
}

// Returns 0 / 1 / 2 (see the header); on 1 and 2 the solution is in S.w.  S.A and S.g are inputs.
template <int K, class EX>
IRLOSC_HD int coop_resolve_pinv(CoopSmem<K> &S, EX &ex) {
    constexpr int MS = 6;
    // ---- factor A, invert, traces
    ex.par([&](int lane) {
        for (int e = lane; e < K * K; e += 32) S.B[e / K][e % K] = S.A[e / K][e % K];
    });
    if (!coop_factor<K>(S, ex)) return 0;
    coop_inverse<K>(S, ex);
    double tr = 0.0, tr_inv = 0.0;
    for (int i = 0; i < K; ++i) { tr += S.A[i][i]; tr_inv += S.Y[i][i]; }
    if (!(tr > 0.0) || !(tr_inv > 0.0)) return 0;
    // ---- bounds on lambda_max, one squaring at a time; outcome 1 needs nothing else
    double sm[MS];
    const double itr = 1.0 / tr;
    ex.par([&](int lane) {
        for (int e = lane; e < K * K; e += 32) S.C[e / K][e % K] = S.A[e / K][e % K] * itr;      // B_m lives in C
    });
    double c_hi = 0.0, c_lo = 0.0;
    for (int m = 0; m <= MS; ++m) {
        if (m > 0) {                                            // C <- C^2 / tr(C^2) via the scratch B
            ex.par([&](int lane) {
                for (int e = lane; e < K * K; e += 32) {
                    const int i = e / K, j = e % K;
                    double acc = 0.0;
                    for (int l = 0; l < K; ++l) acc = fma(S.C[i][l], S.C[l][j], acc);
                    S.B[i][j] = acc;
                }
            });
            double t = 0.0;
            for (int i = 0; i < K; ++i) t += S.B[i][i];
            if (!(t > 0.0)) return 0;
            sm[m - 1] = t;
            const double it2 = 1.0 / t;
            ex.par([&](int lane) {
                for (int e = lane; e < K * K; e += 32) S.C[e / K][e % K] = S.B[e / K][e % K] * it2;
            });
        }
        double up = 1.0, lo = 1.0 / K;                          // lambda_max(B_m) in [1 / K, 1] (PSD, trace 1)
        for (int j = m - 1; j >= 0; --j) { up = sqrt(sm[j] * up); lo = sqrt(sm[j] * lo); }
        c_hi = kPinvRcond * tr * up * (1.0 + 1e-12);
        c_lo = kPinvRcond * tr * lo * (1.0 - 1e-12);
        if (1.0 > c_hi * tr_inv) {                              // outcome 1: w = A^-1 g
            ex.par([&](int lane) {
                if (lane < K) {
                    double acc = 0.0;
                    for (int j = 0; j < K; ++j) acc = fma(S.Y[lane][j], S.g[j], acc);
                    S.w[lane] = acc;
                }
            });
            return 1;
        }
    }
    // ---- not certified with the tightest bounds: smallest eigenpair by inverse iteration (x <- A^-1 x),
    //      until the extrapolated error of x is at rounding level (the contraction lambda_min / lambda_2 is
    //      not bounded away from 1 by the classification)
    ex.par([&](int lane) {
        if (lane < K) S.x[lane] = 1.0 + 0.1 * lane;
    });
    bool converged = false;
    double d_prev = HUGE_VAL;
    for (int it = 0; it < 64 && !converged; ++it) {
        ex.par([&](int lane) {
            if (lane < K) {
                double acc = 0.0;
                for (int j = 0; j < K; ++j) acc = fma(S.Y[lane][j], S.x[j], acc);
                S.xn[lane] = acc;
            }
        });
        double n2 = 0.0, dot = 0.0;
        for (int i = 0; i < K; ++i) { n2 = fma(S.xn[i], S.xn[i], n2); dot = fma(S.xn[i], S.x[i], dot); }
        if (!(n2 > 0.0) || !(n2 < HUGE_VAL)) break;
        const double in = (dot < 0.0 ? -1.0 : 1.0) / sqrt(n2);
        double d2 = 0.0;
        for (int i = 0; i < K; ++i) { const double dl = S.xn[i] * in - S.x[i]; d2 = fma(dl, dl, d2); }
        ex.par([&](int lane) {
            if (lane < K) S.x[lane] = S.xn[lane] * in;
        });
        const double d = sqrt(d2);
        if (it >= 2) {
            const double r = d / d_prev;
            converged = (d == 0.0) || (r < 0.97 && d * r / (1.0 - r) < 1e-13);
        }
        d_prev = d;
    }
    if (!converged) return 0;
    ex.par([&](int lane) {
        if (lane < K) {
            double acc = 0.0;
            for (int j = 0; j < K; ++j) acc = fma(S.A[lane][j], S.x[j], acc);
            S.t[lane] = acc;
        }
    });
    double rho = 0.0;                                           // Rayleigh quotient: lambda_min <= rho
    for (int i = 0; i < K; ++i) rho = fma(S.x[i], S.t[i], rho);
    if (!(rho <= c_lo)) return 0;
    // ---- A' = A + tr(A) x x^T: every other eigenvalue must be above the cutoff
    ex.par([&](int lane) {
        for (int e = lane; e < K * K; e += 32) {
            const int i = e / K, j = e % K;
            S.B[i][j] = fma(tr * S.x[i], S.x[j], S.A[i][j]);
        }
    });
    if (!coop_factor<K>(S, ex)) return 0;
    coop_inverse<K>(S, ex);
    double tr2 = 0.0;
    for (int i = 0; i < K; ++i) tr2 += S.Y[i][i];
    if (!(tr2 > 0.0) || !(1.0 > c_hi * tr2)) return 0;
    double xg = 0.0;                                            // outcome 2: w = A'^-1 (g - x (x . g))
    for (int i = 0; i < K; ++i) xg = fma(S.x[i], S.g[i], xg);
    ex.par([&](int lane) {
        if (lane < K) S.t[lane] = fma(-xg, S.x[lane], S.g[lane]);
    });
    ex.par([&](int lane) {
        if (lane < K) {
            double acc = 0.0;
            for (int j = 0; j < K; ++j) acc = fma(S.Y[lane][j], S.t[j], acc);
            S.w[lane] = acc;
        }
    });
    return 2;
}

}  // namespace fused
}  // namespace irlosc
