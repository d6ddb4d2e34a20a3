// k3_device.cuh -- device helpers shared by the uniformization kernels (k3_unif.cu) and the
// untyped-tree particle kernel (k4_untyped.cu): separately rounded arithmetic, small-matrix
// product, the jblas expm restatement and the branch path sampler of
// operators/UniformizationRetypeOperator.java:74-194.
#pragma once
#include "mtt_common.cuh"

#include <math.h>

namespace {

constexpr int U_WARPS = 4;          // warps per CTA
constexpr int U_TCAP = 512;         // virtual events kept in shared memory per warp
constexpr int U_LOGN = 4096;        // entries of the log(n) table

__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
// one copy each of the long libm / division sequences (instruction-cache footprint)
__device__ __noinline__ double log_ni(double x) { return log(x); }
__device__ __noinline__ double exp_ni(double x) { return exp(x); }
__device__ __noinline__ double div_ni(double a, double b) { return __ddiv_rn(a, b); }

// (row << 8 | col) of every element index of a d x d matrix: filled once per warp so that the
// hot loops need no integer division
__device__ __forceinline__ void fill_rc(uint16_t *rc, int d, int lane)
{
    #pragma unroll 1
    for (int e = lane; e < d * d; e += 32)
        rc[e] = (uint16_t)(((e / d) << 8) | (e % d));
    __syncwarp();
}
#define RC_ROW(v) ((int)(v) >> 8)
#define RC_COL(v) ((int)(v) & 255)

// C = A * B (d x d, row-major, shared memory); C must not alias A or B.  Deliberately NOT
// inlined: expm calls it ~10 times and the kernels must stay instruction-cache resident
// (the fully inlined expm was 4400 SASS instructions and stalled on instruction fetch).
template <bool ILP>
__device__ __noinline__ void mm(int d, const double *A, const double *B, double *C,
                                const uint16_t *rc, int lane)
{
    if (d == 2) {
        // two demes (the common case): four lanes, two products each, same rounding sequence
        if (lane < 4) {
            const int i = lane >> 1, j = lane & 1;
            double s = add_rn(0.0, mul_rn(A[2 * i], B[j]));
            C[lane] = add_rn(s, mul_rn(A[2 * i + 1], B[2 + j]));
        }
        __syncwarp();
        return;
    }
    const int dd = d * d;
    if (ILP) {
        // four elements of C per lane and pass, as four independent sums: the rounding
        // sequence of every element is the sequential one, only the dependent chains overlap
        // (the latency-bound kernels of a few dozen chains; costs registers)
        #pragma unroll 1
        for (int e0 = lane; e0 < dd; e0 += 128) {
            int ia[4], jb[4];
            double s[4];
            #pragma unroll
            for (int q = 0; q < 4; q++) {
                const int e = min(e0 + 32 * q, dd - 1);     // (clamped: computed, not stored)
                ia[q] = RC_ROW(rc[e]) * d;
                jb[q] = RC_COL(rc[e]);
                s[q] = 0.0;
            }
            for (int k = 0; k < d; k++) {
                #pragma unroll
                for (int q = 0; q < 4; q++)
                    s[q] = add_rn(s[q], mul_rn(A[ia[q] + k], B[k * d + jb[q]]));
            }
            #pragma unroll
            for (int q = 0; q < 4; q++)
                if (e0 + 32 * q < dd)
                    C[e0 + 32 * q] = s[q];
        }
    } else {
        #pragma unroll 1
        for (int e = lane; e < dd; e += 32) {
            const int i = RC_ROW(rc[e]), j = RC_COL(rc[e]);
            double s = 0.0;
            for (int k = 0; k < d; k++)
                s = add_rn(s, mul_rn(A[i * d + k], B[k * d + j]));
            C[e] = s;
        }
    }
    __syncwarp();
}

// ---- jblas MatrixFunctions.expm restated [EXT]: Pade-13 scaling & squaring + LU solve ------
__constant__ double PADE[14] = {
    1.0, 0.5, 0.12, 0.01833333333333333, 0.0019927536231884053, 1.630434782608695E-4,
    1.0351966873706E-5, 5.175983436853E-7, 2.0431513566525E-8, 6.306022705717593E-10,
    1.4837700484041396E-11, 2.5291534915979653E-13, 2.8101705462199615E-15,
    1.5440497506703084E-17};

__device__ __forceinline__ double pade_poly(double eye, double a2, double a4, double a6, int c0)
{
    // ((eye*c[c0] + A2*c[c0+2]) + A4*c[c0+4]) + A6*c[c0+6]
    return add_rn(add_rn(add_rn(mul_rn(eye, PADE[c0]), mul_rn(a2, PADE[c0 + 2])),
                         mul_rn(a4, PADE[c0 + 4])),
                  mul_rn(a6, PADE[c0 + 6]));
}

// LU with partial pivoting, F = D^-1 N left in N (unblocked right-looking elimination)
__device__ __noinline__ void lu_solve(int d, double *D, double *N, const uint16_t *rc, int lane)
{
    for (int k = 0; k < d; k++) {
        int p = k;
        double best = fabs(D[k * d + k]);
        for (int i = k + 1; i < d; i++) {
            const double v = fabs(D[i * d + k]);
            if (v > best) {
                best = v;
                p = i;
            }
        }
        __syncwarp();
        if (p != k) {
            #pragma unroll 1
            for (int j = lane; j < d; j += 32) {
                double t = D[k * d + j]; D[k * d + j] = D[p * d + j]; D[p * d + j] = t;
                t = N[k * d + j]; N[k * d + j] = N[p * d + j]; N[p * d + j] = t;
            }
            __syncwarp();
        }
        const double pivot = D[k * d + k];
        #pragma unroll 1
        for (int i = k + 1 + lane; i < d; i += 32)
            D[i * d + k] = div_ni(D[i * d + k], pivot);
        __syncwarp();
        const int rows = d - k - 1;
        #pragma unroll 1
        for (int e = lane; e < rows * d; e += 32) {
            const int i = k + 1 + RC_ROW(rc[e]), j = RC_COL(rc[e]);
            const double l = D[i * d + k];
            if (j > k)
                D[i * d + j] = sub_rn(D[i * d + j], mul_rn(l, D[k * d + j]));
            N[i * d + j] = sub_rn(N[i * d + j], mul_rn(l, N[k * d + j]));
        }
        __syncwarp();
    }
    for (int k = d - 1; k >= 0; k--) {
        const double pivot = D[k * d + k];
        __syncwarp();
        #pragma unroll 1
        for (int j = lane; j < d; j += 32)
            N[k * d + j] = div_ni(N[k * d + j], pivot);
        __syncwarp();
        #pragma unroll 1
        for (int e = lane; e < k * d; e += 32) {
            const int i = RC_ROW(rc[e]), j = RC_COL(rc[e]);
            N[i * d + j] = sub_rn(N[i * d + j], mul_rn(D[i * d + k], N[k * d + j]));
        }
        __syncwarp();
    }
}

// expm(Q * L) with Q in global/shared memory; `buf` = 8 d*d shared-memory matrices of this
// warp.  Returns a pointer (inside buf) to the result.
template <bool ILP = false>
__device__ __noinline__ const double *expm_QL(int d, const double *Q, double L, double *buf,
                                              const uint16_t *rc, int lane)
{
    const int dd = d * d;
    double *As = buf, *A2 = buf + dd, *A4 = buf + 2 * dd, *A6 = buf + 3 * dd, *T = buf + 4 * dd,
           *U = buf + 5 * dd, *V = buf + 6 * dd, *W = buf + 7 * dd;
    double nm = 0.0;
    #pragma unroll 1
    for (int e = lane; e < dd; e += 32) {
        const double a = mul_rn(Q[e], L);                   // Q.mul(L)
        As[e] = a;
        nm = fmax(nm, fabs(a));
    }
    nm = warp_max(nm);
    int j = 0;
    if (nm > 0.0) {
        // log(2.0) is the correctly rounded constant on host and device alike
        const double f = floor(log_ni(nm) / 0.6931471805599453);
        if (f > -1.0)
            j = 1 + (int)f;
    }
    const double invScale = scalbn(1.0, -j);     // dividing by 2^j == multiplying by 2^-j, exactly
    __syncwarp();
    #pragma unroll 1
    for (int e = lane; e < dd; e += 32)
        As[e] = mul_rn(As[e], invScale);
    __syncwarp();
    mm<ILP>(d, As, As, A2, rc, lane);
    mm<ILP>(d, A2, A2, A4, rc, lane);
    mm<ILP>(d, A4, A2, A6, rc, lane);
    #pragma unroll 1
    for (int e = lane; e < dd; e += 32)
        T[e] = pade_poly((RC_ROW(rc[e]) == RC_COL(rc[e])) ? 1.0 : 0.0, A2[e], A4[e], A6[e], 6);
    __syncwarp();
    mm<ILP>(d, T, A6, U, rc, lane);
    #pragma unroll 1
    for (int e = lane; e < dd; e += 32) {
        const double eye = (RC_ROW(rc[e]) == RC_COL(rc[e])) ? 1.0 : 0.0;
        U[e] = add_rn(add_rn(add_rn(mul_rn(eye, PADE[0]), mul_rn(A2[e], PADE[2])),
                             mul_rn(A4[e], PADE[4])),
                      U[e]);
    }
    __syncwarp();
    #pragma unroll 1
    for (int e = lane; e < dd; e += 32)
        T[e] = pade_poly((RC_ROW(rc[e]) == RC_COL(rc[e])) ? 1.0 : 0.0, A2[e], A4[e], A6[e], 7);
    __syncwarp();
    mm<ILP>(d, T, A6, V, rc, lane);
    #pragma unroll 1
    for (int e = lane; e < dd; e += 32) {
        const double eye = (RC_ROW(rc[e]) == RC_COL(rc[e])) ? 1.0 : 0.0;
        V[e] = add_rn(add_rn(add_rn(mul_rn(eye, PADE[1]), mul_rn(A2[e], PADE[3])),
                             mul_rn(A4[e], PADE[5])),
                      V[e]);
    }
    __syncwarp();
    mm<ILP>(d, As, V, W, rc, lane);                                  // AV
    double *Nn = A2, *D = A4;                               // A2, A4 are dead from here
    #pragma unroll 1
    for (int e = lane; e < dd; e += 32) {
        const double u = U[e], av = W[e];
        Nn[e] = add_rn(u, av);
        D[e] = sub_rn(u, av);
    }
    __syncwarp();
    lu_solve(d, D, Nn, rc, lane);
    double *F = Nn, *G = T;
    for (int k = 0; k < j; k++) {                           // F = F * F, j times
        mm<ILP>(d, F, F, G, rc, lane);
        double *t = F;
        F = G;
        G = t;
    }
    return F;
}

// ascending bitonic sort of npad (power of two) doubles by one warp
__device__ __forceinline__ void warp_sort_doubles(double *x, int npad, int lane)
{
    for (int size = 2; size <= npad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = lane; i < (npad >> 1); i += 32) {
                const int a = 2 * i - (i & (stride - 1));
                int b = a + stride;
                if (stride == (size >> 1))
                    b = (a & ~(size - 1)) + (size - 1 - (a & (size - 1)));
                const double xa = x[a], xb = x[b];
                if (xa > xb) {
                    x[a] = xb;
                    x[b] = xa;
                }
            }
            __syncwarp();
        }
    }
}

// Poisson(lam) from the uniform window (multiplicative method below 40, inversion by summation
// above; the documented stand-in for BEAST's [EXT] Randomizer.nextPoisson).  false = window
// exhausted.
__device__ __noinline__ bool poisson_window(double lam, const double *u, long long nU,
                                            long long &pos, int &out)
{
    if (lam < 40.0) {
        const double pz = exp_ni(-lam);
        double r = 1.0;
        int n = 0;
        for (;;) {
            if (pos >= nU)
                return false;
            r = mul_rn(r, u[pos++]);
            if (r < pz) {
                out = n;
                return true;
            }
            n++;
        }
    }
    if (pos >= nU)
        return false;
    const double uu = u[pos++];
    int n = 0;
    double logp = -lam, cdf = exp_ni(logp);
    const double loglam = log_ni(lam);
    while (cdf < uu && (double)n < add_rn(mul_rn(10.0, lam), 100.0)) {
        n++;
        logp = add_rn(logp, sub_rn(loglam, log_ni((double)n)));
        cdf = add_rn(cdf, exp_ni(logp));
    }
    out = n;
    return true;
}

struct BranchSample {
    int st, nVirt;
    double Pba, logPba;  // logPba: NaN unless the event count was drawn (then log(Pba))
    double *times;       // sorted times of the virtual events
    uint8_t *types;      // type after each virtual event
};

// retypeBranch up to and including the forward-backward pass (:124-194), by one warp.  On
// MTT_ST_OK the nVirt sorted virtual-event times and sampled types are in .times / .types
// (shared memory, or the caller's global scratch when nVirt > U_TCAP).  tab / size / steady
// describe the model's R^k table, logTab the elementwise log of it, logN[n] = log(n).
// TAIL_ON_DEVICE draws the steady-state Poisson tail
// (:96-102) from the window instead of handing it to the host.
// LONG_BRANCHES adds the paths that cut the LATENCY of a branch with many virtual events (the
// retype kernel of a few dozen chains waits for its slowest warp); they cost registers, so the
// throughput-bound particle kernel leaves them out.  Results are identical either way.
template <bool TAIL_ON_DEVICE, bool LONG_BRANCHES = false>
__device__ __forceinline__ BranchSample sample_branch(
    int d, const double *Q, const double *R, const double *tab, const double *logTab,
    const double *logN, int size, int steady, double mu,
    double t0, double t1, int a, int b, int forced, const double *u, long long nU,
    long long &pos, double *buf, const uint16_t *rc, double *sTimes, uint8_t *sTypes,
    double *gTimes, uint8_t *gTypes, int scratchCap, int lane, int stageCap = U_TCAP,
    const double *PbaPre = nullptr)
{
    const int dd = d * d;
    const int steadyN = steady ? size : -1;                          // RpowSteadyN :460-473
    const double L = sub_rn(t1, t0);                                 // :124
    const double muL = mul_rn(mu, L);                         // :130
    // expm(Q L)[a,b] (:132-134): evaluated here, or ahead of the sequential sampling by
    // another warp running the same expm_QL (k3_pba) -- the same bits either way
    double Pba;
    if (PbaPre) {
        Pba = *PbaPre;
    } else {
        const double *E = expm_QL<LONG_BRANCHES>(d, Q, L, buf, rc, lane);
        Pba = E[a * d + b];
    }
    __syncwarp();

    int st = MTT_ST_OK, nVirt = -1;
    double logPba = NAN;
    if (Pba == 0.0)                                                  // :137-138
        st = MTT_ST_NO_VALID_PATH;
    else if (Pba > 1.0 || Pba < 0.0)                                 // :141-144
        st = MTT_ST_PBA_RUBBISH;

    // ---- drawEventCount (:74-105)
    // The loop is a chain of dependent operations; the table entries of step n + 1 (log R^n[a,b],
    // log n) do not depend on it and are fetched one step ahead.
    if (st == MTT_ST_OK) {
        if (forced >= 0) {
            nVirt = forced;
        } else if (pos >= nU) {
            st = MTT_ST_NEED_MORE_UNIFORMS;
        } else {
            const double uu = u[pos++];                              // :79
            double Plow = 0.0;
            logPba = log_ni(Pba);
            double acc = sub_rn(-muL, logPba);                          // :81
            const double log_muL = log_ni(muL);                         // :82
            const int ab = a * d + b;
            int n = 0;
            double curLogRn = (size > 0) ? logTab[ab] : 0.0;         // log R^0[a,b]
            double curLogN = logN[1];                                // log(n) of the next step
            // what the reference does when the table turned steady before the draw was found
            auto steadyTail = [&]() {                                // :96-102
                st = MTT_ST_NEEDS_POISSON_TAIL;                      // -> host
                nVirt = n;                                           // thresh
                if (TAIL_ON_DEVICE) {
                    // rejection from Poisson(muL) until >= thresh, from the same window
                    for (;;) {
                        int draw = 0;
                        if (!poisson_window(muL, u, nU, pos, draw)) {
                            st = MTT_ST_NEED_MORE_UNIFORMS;
                            break;
                        }
                        if (draw >= n) {
                            nVirt = draw;
                            st = MTT_ST_OK;
                            break;
                        }
                    }
                }
            };
            // (1) the first terms one by one: most branches are short and stop here
            constexpr int SEQ_TERMS = LONG_BRANCHES ? 6 : 0x7fffffff;
            bool done = false;
            while (!done && n < SEQ_TERMS) {
                if (n >= size && !steady) {
                    st = MTT_ST_NEED_MORE_POWERS;
                    done = true;
                    break;
                }
                const int n1 = n + 1;
                double nxtLogRn = 0.0, nxtLogN = 0.0;
                if (n1 < size || steady)
                    nxtLogRn = logTab[(size_t)min(n1, size - 1) * dd + ab];
                if (n1 + 1 < U_LOGN)
                    nxtLogN = logN[n1 + 1];
                // log(R^n[a,b]) and log(n) come from tables filled with the same log()
                Plow = add_rn(Plow, exp_ni(add_rn(curLogRn, acc)));        // :86
                if (Plow > uu) {                                     // :88-89
                    nVirt = n;
                    done = true;
                    break;
                }
                n += 1;
                acc = add_rn(acc, sub_rn(log_muL, n < U_LOGN ? curLogN : log_ni((double)n)));  // :92
                curLogRn = nxtLogRn;
                curLogN = nxtLogN;
                if (!(steadyN < 0 || n < steadyN)) {                 // :94
                    steadyTail();
                    done = true;
                }
            }
            // (2) a long branch: 32 terms at a time.  Lane j owns term n + j; the running
            // exponent acc_(n+j) is the reference's chain of rounded additions (cheap, done by
            // every lane), the exp() of the 32 terms is evaluated side by side, and the partial
            // sums are then added up and tested in the reference's order.
            while (LONG_BRANCHES && !done) {
                const int nj = n + lane;
                const int m = nj + 1;                                // index of the log(n) the step after term nj uses
                const double dj = sub_rn(log_muL, m < U_LOGN ? logN[m] : log_ni((double)m));
                const double logRj = logTab[(size_t)min(nj, size - 1) * dd + ab];
                double run = acc, mine = acc;
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    if (lane == j)
                        mine = run;
                    run = add_rn(run, __shfl_sync(FULL_MASK, dj, j));
                }
                const double term = exp_ni(add_rn(logRj, mine));
                for (int j = 0; j < 32; j++) {
                    if (n >= size && !steady) {
                        st = MTT_ST_NEED_MORE_POWERS;
                        done = true;
                        break;
                    }
                    Plow = add_rn(Plow, __shfl_sync(FULL_MASK, term, j));   // :86
                    if (Plow > uu) {                                 // :88-89
                        nVirt = n;
                        done = true;
                        break;
                    }
                    n += 1;                                          // :92 (acc: `run`, below)
                    if (!(steadyN < 0 || n < steadyN)) {             // :94
                        steadyTail();
                        done = true;
                        break;
                    }
                }
                acc = run;
            }
        }
    }
    // a steady table answers any n; otherwise powers up to nVirt must exist
    if (st == MTT_ST_OK && nVirt >= size && !steady)
        st = MTT_ST_NEED_MORE_POWERS;
    if (st == MTT_ST_OK && pos + 2LL * nVirt > nU)
        st = MTT_ST_NEED_MORE_UNIFORMS;
    double *times = sTimes;
    uint8_t *types = sTypes;
    if (st == MTT_ST_OK && nVirt > stageCap) {
        if (nVirt > scratchCap) {
            st = MTT_ST_NEED_MORE_POWERS;   // host grows the scratch along with the table
        } else {
            times = gTimes;
            types = gTypes;
        }
    }

    if (st == MTT_ST_OK) {
        // ---- times of the virtual events (:154-157)
        int npad = 1;
        while (npad < nVirt)
            npad <<= 1;
        if (nVirt == 0)
            npad = 0;
        for (int i = lane; i < npad; i += 32)
            times[i] = (i < nVirt) ? add_rn(mul_rn(u[pos + i], L), t0) : INFINITY;
        pos += nVirt;
        __syncwarp();
        warp_sort_doubles(times, npad, lane);

        // ---- forward-backward sampling of the types (:163-194)
        // Step i needs column b of R^(nVirt-i+1) and of R^(nVirt-i): lane c holds R^k[c,b].  The
        // first is the previous step's second, and the next step's column does not depend on
        // the type drawn here, so it is fetched one step ahead; R itself sits in the warp's
        // shared scratch (free once Pba is known).
        for (int e = lane; e < dd; e += 32)
            buf[e] = R[e];
        __syncwarp();
        auto column = [&](int k) {
            return lane < d ? tab[(size_t)min(k, size - 1) * dd + lane * d + b] : 0.0;
        };
        double hiCol = nVirt >= 1 ? column(nVirt) : 0.0;
        double loCol = nVirt >= 1 ? column(nVirt - 1) : 0.0;
        int prev = a;
        for (int i = 1; i <= nVirt; i++) {
            const double nxtCol = (i < nVirt) ? column(nVirt - i - 1) : 0.0;
            double u2 = mul_rn(u[pos++], __shfl_sync(FULL_MASK, hiCol, prev));   // :165-166
            double prod = 0.0;
            if (lane < d)
                prod = mul_rn(buf[prev * d + lane], loCol);
            int c = d;
            if (LONG_BRANCHES && d > 4) {
                // many demes: the d products first (independent shuffles), then the chain of
                // subtractions in the reference's order, stopping at the first negative value
                double pr[MTT_MAX_DEMES];
#pragma unroll
                for (int q = 0; q < MTT_MAX_DEMES; q++)
                    pr[q] = __shfl_sync(FULL_MASK, prod, q);
#pragma unroll
                for (int q = 0; q < MTT_MAX_DEMES; q++) {
                    if (q < d && c == d) {
                        u2 = sub_rn(u2, pr[q]);                      // :170-171
                        if (u2 < 0.0)
                            c = q;
                    }
                }
            } else {
                for (int q = 0; q < d; q++) {
                    u2 = sub_rn(u2, __shfl_sync(FULL_MASK, prod, q));    // :170-171
                    if (u2 < 0.0) {
                        c = q;
                        break;
                    }
                }
            }
            const bool fell = c == d;
            if (fell) {                                              // :179-190
                st = MTT_ST_FB_FELL_THROUGH;
                break;
            }
            if (lane == 0)
                types[i - 1] = (uint8_t)c;
            prev = c;
            hiCol = loCol;
            loCol = nxtCol;
        }
        __syncwarp();
    }
    BranchSample r;
    r.st = st;
    r.nVirt = nVirt;
    r.Pba = Pba;
    r.logPba = logPba;
    r.times = times;
    r.types = types;
    return r;
}

}  // namespace
