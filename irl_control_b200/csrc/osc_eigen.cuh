// Warp-cooperative symmetric eigen-solver shared by the step kernels (pinv / indefinite fallback
// of osc.py:52-55) and the fix-up kernel of the fused step.
#pragma once
#include "irlosc_device.cuh"

namespace irlosc {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

namespace tiled {

// Warp-cooperative symmetric eigen-solve for one instance (pinv / indefinite fallback):
// parallel-order Jacobi.  A round-robin schedule gives K/2 disjoint (p, q) pairs per round, so one
// round applies K/2 rotations with three passes over the matrix (columns of A and V, rows of A,
// clean-up) instead of one barrier-separated pass per rotation.
// A (destroyed, leading dimension LDA) and V (K+1); cbuf / sbuf hold >= K doubles each;
// result w = V f(lambda) V^T g.
template <int K, int LDA = K + 1>
__device__ void eigen_solve(double (*A)[LDA], double (*V)[K + 1], const double *g, double *w, double *cbuf,
                            double *sbuf, bool force_pinv, int lane, int *flags_out) {
    constexpr int M = (K % 2 == 0) ? K : K + 1;      // players of the round-robin (one dummy when K is odd)
    constexpr int KP = M / 2;                        // pairs per round
    auto pair_of = [](int r, int t, int &p, int &q) {
        int a, b;
        if (t == 0) { a = M - 1; b = r % (M - 1); }
        else { a = (r + t) % (M - 1); b = (r - t + 2 * (M - 1)) % (M - 1); }
        p = a < b ? a : b;
        q = a < b ? b : a;
    };
    for (int i = lane; i < K * K; i += 32) V[i / K][i % K] = (i / K == i % K) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, dia = 0.0;
        for (int i = lane; i < K * K; i += 32) {
            const double v = A[i / K][i % K];
            if (i / K == i % K) dia += v * v; else off += v * v;
        }
        off = warp_sum(off);
        dia = warp_sum(dia);
        if (off <= 1e-28 * dia || off == 0.0) break;
        for (int r = 0; r < M - 1; ++r) {
            if (lane < KP) {                         // rotation of pair `lane`
                int p, q;
                pair_of(r, lane, p, q);
                double c = 1.0, sn = 0.0;
                if (q < K) {
                    const double apq = A[p][q], app = A[p][p], aqq = A[q][q];
                    if (!(fabs(apq) <= 1e-300 || apq * apq <= 1e-31 * fabs(app * aqq))) {
                        const double theta = (aqq - app) * fast_rcp(2.0 * apq);
                        const double tt = (theta >= 0.0 ? 1.0 : -1.0) * fast_rcp(fabs(theta) + fast_sqrt(theta * theta + 1.0));
                        c = fast_rcp(fast_sqrt(tt * tt + 1.0));
                        sn = tt * c;
                    }
                }
                cbuf[lane] = c;
                sbuf[lane] = sn;
            }
            __syncwarp();
            for (int e = lane; e < K * KP; e += 32) {        // A <- A J, V <- V J (columns p, q)
                const int i = e / KP, t = e % KP;
                int p, q;
                pair_of(r, t, p, q);
                if (q < K) {
                    const double c = cbuf[t], sn = sbuf[t];
                    const double aip = A[i][p], aiq = A[i][q];
                    A[i][p] = c * aip - sn * aiq;
                    A[i][q] = sn * aip + c * aiq;
                    const double vip = V[i][p], viq = V[i][q];
                    V[i][p] = c * vip - sn * viq;
                    V[i][q] = sn * vip + c * viq;
                }
            }
            __syncwarp();
            for (int e = lane; e < K * KP; e += 32) {        // A <- J^T A (rows p, q)
                const int j = e / KP, t = e % KP;
                int p, q;
                pair_of(r, t, p, q);
                if (q < K) {
                    const double c = cbuf[t], sn = sbuf[t];
                    const double apj = A[p][j], aqj = A[q][j];
                    A[p][j] = c * apj - sn * aqj;
                    A[q][j] = sn * apj + c * aqj;
                }
            }
            __syncwarp();
            if (lane < KP) {                         // the rotated pair is exactly decoupled
                int p, q;
                pair_of(r, lane, p, q);
                if (q < K && sbuf[lane] != 0.0) { A[p][q] = 0.0; A[q][p] = 0.0; }
            }
            __syncwarp();
        }
    }
    __syncwarp();
    double lmax = 0.0, det = 1.0;
    for (int i = 0; i < K; ++i) { lmax = fmax(lmax, fabs(A[i][i])); det *= A[i][i]; }
    const bool pinv = force_pinv || !(fabs(det) >= kDetThreshold);
    if (lane < K) {
        const double lam = A[lane][lane];
        double proj = 0.0;
        for (int i = 0; i < K; ++i) proj += V[i][lane] * g[i];
        const bool keep = pinv ? (fabs(lam) > kPinvRcond * lmax) : true;
        cbuf[lane] = keep ? proj / lam : 0.0;
    }
    __syncwarp();
    if (lane < K) {
        double acc = 0.0;
        for (int c = 0; c < K; ++c) acc += V[lane][c] * cbuf[c];
        w[lane] = acc;
    }
    if (lane == 0) *flags_out |= IRLOSC_ST_EIGEN | (pinv ? IRLOSC_ST_PINV : 0);
    __syncwarp();
}

}  // namespace tiled
}  // namespace irlosc
