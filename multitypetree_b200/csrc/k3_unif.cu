// k3_unif.cu -- piece 3: uniformization kernels on sm_100a.
//
// Replaces evolution/tree/SCMigrationModel.java:150-201 (updateMatrices), :404-473 (getRpowN /
// RpowSteadyN -- the jblas mmul chain), and operators/UniformizationRetypeOperator.java:74-275
// (drawEventCount, retypeBranch, getBranchTypeProb -- jblas expm + forward-backward sampling).
//
// Deme counts are tiny (d <= 16): no tensor cores.  Each model / branch / job is one warp;
// d x d matrices live in shared memory (a lane owns elements lane, lane+32, ...), products
// are the naive triple loop.  Arithmetic contract: every product / sum that feeds R^k, expm
// or a sampling comparison uses separately rounded multiply and add (__dmul_rn / __dadd_rn,
// which the compiler never contracts into FMA) in the reference's association order, so
// R^k, expm(QL) and the sampled type paths are bit-identical to the CPU restatement for the
// same uniform stream (only libm log/exp may differ in the last ulp).
#include "mtt_common.cuh"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace {

constexpr int U_WARPS = 4;          // warps per CTA
constexpr int U_TCAP = 512;         // virtual events kept in shared memory per warp

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
__device__ __noinline__ void mm(int d, const double *A, const double *B, double *C,
                                const uint16_t *rc, int lane)
{
    #pragma unroll 1
    for (int e = lane; e < d * d; e += 32) {
        const int i = RC_ROW(rc[e]), j = RC_COL(rc[e]);
        double s = 0.0;
        for (int k = 0; k < d; k++)
            s = add_rn(s, mul_rn(A[i * d + k], B[k * d + j]));
        C[e] = s;
    }
    __syncwarp();
}

// ---- updateMatrices (SCMigrationModel.java:150-201), one warp per model ------------------
__global__ void k3_update_matrices(int nModels, int d, const double *m, int sym, double *Q,
                                   double *R, double *mu)
{
    const int model = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (model >= nModels)
        return;
    const double *mm_ = m + (size_t)model * d * d;
    double *q = Q + (size_t)model * d * d, *r = R + (size_t)model * d * d;
    double diagNeg = 0.0;
    for (int i = lane; i < d; i += 32) {
        double qii = 0.0;
        for (int j = 0; j < d; j++) {
            if (i != j) {
                double v = sym ? mul_rn(0.5, add_rn(mm_[i * d + j], mm_[j * d + i])) : mm_[i * d + j];
                q[i * d + j] = v;
                qii = sub_rn(qii, v);                       // :166-168 / :170-171
            }
        }
        q[i * d + i] = qii;
        diagNeg = fmax(diagNeg, -qii);                      // :175-179 (mu starts at 0.0)
    }
    const double muV = warp_max(diagNeg);
    __syncwarp();
    const double rm = div_ni(1.0, muV);
    #pragma unroll 1
    for (int e = lane; e < d * d; e += 32) {
        const int i = e / d, j = e - i * d;
        r[e] = add_rn(mul_rn(q[e], rm), (i == j) ? 1.0 : 0.0);   // Q.mul(1.0/mu).add(eye) :183
    }
    if (lane == 0)
        mu[model] = muV;
}

// ---- R^k chain (SCMigrationModel.java:404-452), one warp per model -------------------------
// meta[2*model] = number of valid powers, meta[2*model+1] = steady flag
__global__ void k3_rpow_extend(int nModels, int d, const double *R, double *pow, int powCap,
                               int *meta, int target)
{
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int model = blockIdx.x * U_WARPS + warp;
    if (model >= nModels)
        return;
    const int dd = d * d;
    __shared__ uint16_t rcAll[U_WARPS][MTT_MAX_DEMES * MTT_MAX_DEMES];
    uint16_t *rc = rcAll[threadIdx.x >> 5];
    fill_rc(rc, d, threadIdx.x & 31);
    double *sR = sm + (size_t)warp * 3 * dd, *sPrev = sR + dd, *sNew = sPrev + dd;
    double *tab = pow + (size_t)model * powCap * dd;
    int size = meta[2 * model], steady = meta[2 * model + 1];
    #pragma unroll 1
    for (int e = lane; e < dd; e += 32)
        sR[e] = R[(size_t)model * dd + e];
    if (size == 0) {                                        // RpowN = [I]  (:193)
        #pragma unroll 1
        for (int e = lane; e < dd; e += 32)
            tab[e] = (RC_ROW(rc[e]) == RC_COL(rc[e])) ? 1.0 : 0.0;
        size = 1;
    }
    #pragma unroll 1
    for (int e = lane; e < dd; e += 32)
        sPrev[e] = tab[(size_t)(size - 1) * dd + e];
    __syncwarp();
    if (target >= powCap)
        target = powCap - 1;
    for (int i = size; i <= target && !steady; i++) {       // :428-448
        mm(d, sPrev, sR, sNew, rc, lane);
        bool anyNaN = false, anyPos = false;
        #pragma unroll 1
        for (int e = lane; e < dd; e += 32) {
            const double v = sNew[e];
            tab[(size_t)i * dd + e] = v;
            const double diff = fabs(v - sPrev[e]);
            anyNaN |= isnan(diff);
            anyPos |= diff > 0.0;
        }
        size = i + 1;
        if (i % 10 == 0) {                                  // :435-447
            anyNaN = __any_sync(FULL_MASK, anyNaN);
            anyPos = __any_sync(FULL_MASK, anyPos);
            if (anyNaN || !anyPos)                          // !(maxDiff > 0)
                steady = 1;
        }
        __syncwarp();
        double *t = sPrev;
        sPrev = sNew;
        sNew = t;
    }
    if (lane == 0) {
        meta[2 * model] = size;
        meta[2 * model + 1] = steady;
    }
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
        const double f = floor(log_ni(nm) / log_ni(2.0));
        if (f > -1.0)
            j = 1 + (int)f;
    }
    const double scale = scalbn(1.0, j);
    __syncwarp();
    #pragma unroll 1
    for (int e = lane; e < dd; e += 32)
        As[e] = div_ni(As[e], scale);
    __syncwarp();
    mm(d, As, As, A2, rc, lane);
    mm(d, A2, A2, A4, rc, lane);
    mm(d, A4, A2, A6, rc, lane);
    #pragma unroll 1
    for (int e = lane; e < dd; e += 32)
        T[e] = pade_poly((RC_ROW(rc[e]) == RC_COL(rc[e])) ? 1.0 : 0.0, A2[e], A4[e], A6[e], 6);
    __syncwarp();
    mm(d, T, A6, U, rc, lane);
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
    mm(d, T, A6, V, rc, lane);
    #pragma unroll 1
    for (int e = lane; e < dd; e += 32) {
        const double eye = (RC_ROW(rc[e]) == RC_COL(rc[e])) ? 1.0 : 0.0;
        V[e] = add_rn(add_rn(add_rn(mul_rn(eye, PADE[1]), mul_rn(A2[e], PADE[3])),
                             mul_rn(A4[e], PADE[5])),
                      V[e]);
    }
    __syncwarp();
    mm(d, As, V, W, rc, lane);                                  // AV
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
        mm(d, F, F, G, rc, lane);
        double *t = F;
        F = G;
        G = t;
    }
    return F;
}

__global__ void k3_expm_batch(long long n, int d, const double *Q, const int *model,
                              const double *L, double *out)
{
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * U_WARPS + warp;
    if (item >= n)
        return;
    const int dd = d * d;
    __shared__ uint16_t rcAll[U_WARPS][MTT_MAX_DEMES * MTT_MAX_DEMES];
    uint16_t *rc = rcAll[threadIdx.x >> 5];
    fill_rc(rc, d, threadIdx.x & 31);
    double *buf = sm + (size_t)warp * 8 * dd;
    const double *F = expm_QL(d, Q + (size_t)model[item] * dd, L[item], buf, rc, lane);
    #pragma unroll 1
    for (int e = lane; e < dd; e += 32)
        out[item * dd + e] = F[e];
}

// ---- getBranchTypeProb (UniformizationRetypeOperator.java:233-275), one warp per branch ----
__global__ void k3_branch_logp(long long n, int d, const double *Qall, const int *model,
                               const double *tBottom, const double *tTop,
                               const uint8_t *typeBottom, const uint8_t *typeTop,
                               const long long *chgOff, const double *chgTime,
                               const uint8_t *chgType, double *logProb, int *status)
{
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long br = (long long)blockIdx.x * U_WARPS + warp;
    if (br >= n)
        return;
    const int dd = d * d;
    __shared__ uint16_t rcAll[U_WARPS][MTT_MAX_DEMES * MTT_MAX_DEMES];
    uint16_t *rc = rcAll[threadIdx.x >> 5];
    fill_rc(rc, d, threadIdx.x & 31);
    double *buf = sm + (size_t)warp * 8 * dd;
    const double *Q = Qall + (size_t)model[br] * dd;
    const double t0 = tBottom[br], t1 = tTop[br];
    const int a = typeBottom[br], b = typeTop[br];
    double lp = 0.0, lastTime = t0;
    int lastCol = a;
    for (long long k = chgOff[br]; k < chgOff[br + 1]; k++) {           // :249-258
        const double thisTime = chgTime[k];
        const int thisCol = chgType[k];
        lp = add_rn(lp, add_rn(mul_rn(sub_rn(thisTime, lastTime), Q[lastCol * d + lastCol]),
                               log_ni(Q[lastCol * d + thisCol])));
        lastTime = thisTime;
        lastCol = thisCol;
    }
    lp = add_rn(lp, mul_rn(sub_rn(t1, lastTime), Q[lastCol * d + lastCol]));   // :259
    const double *E = expm_QL(d, Q, sub_rn(t1, t0), buf, rc, lane);          // :263-264
    const double Pba = E[a * d + b];
    if (lane == 0) {
        if (Pba > 1.0 || Pba < 0.0) {                                    // :267-270
            logProb[br] = -INFINITY;
            status[br] = MTT_ST_PBA_RUBBISH;
        } else {
            logProb[br] = sub_rn(lp, log_ni(Pba));                          // :272
            status[br] = MTT_ST_OK;
        }
    }
}

// ---- retypeBranch + drawEventCount (:74-225), one warp per job ----------------------------
struct RetypeParams {
    long long nJobs;
    int d, powCap, outCap;
    const double *Q, *R, *mu, *pow;
    const int *meta;
    const long long *brOff;
    const int *model;
    const double *tBottom, *tTop;
    const uint8_t *typeBottom, *typeTop;
    const int *nVirtForced;
    const double *uniforms;
    const long long *uOff;
    int *nChg;
    double *outChgTime;
    uint8_t *outChgType;
    double *logProb;
    int *status, *nVirt;
    long long *uUsed;
    double *scratchT;        // [warps][scratchCap] spill for nVirt > U_TCAP
    uint8_t *scratchTy;
    int scratchCap;
};

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

__global__ void __launch_bounds__(U_WARPS * 32) k3_retype(const RetypeParams p)
{
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long job = (long long)blockIdx.x * U_WARPS + warp;
    if (job >= p.nJobs)
        return;
    const int d = p.d, dd = d * d;
    __shared__ uint16_t rcAll[U_WARPS][MTT_MAX_DEMES * MTT_MAX_DEMES];
    uint16_t *rc = rcAll[threadIdx.x >> 5];
    fill_rc(rc, d, threadIdx.x & 31);
    const size_t perWarp = (size_t)8 * dd + U_TCAP + U_TCAP / 8;
    double *buf = sm + (size_t)warp * perWarp;
    double *sTimes = buf + 8 * dd;
    uint8_t *sTypes = reinterpret_cast<uint8_t *>(sTimes + U_TCAP);

    const double *u = p.uniforms + p.uOff[job];
    const long long nU = p.uOff[job + 1] - p.uOff[job];
    long long pos = 0;
    bool aborted = false;

    for (long long br = p.brOff[job]; br < p.brOff[job + 1]; br++) {
        if (aborted) {
            if (lane == 0) {
                p.status[br] = MTT_ST_SKIPPED;
                p.nChg[br] = 0;
                p.logProb[br] = -INFINITY;
                p.nVirt[br] = -1;
            }
            continue;
        }
        const int mdl = p.model[br];
        const double *Q = p.Q + (size_t)mdl * dd, *R = p.R + (size_t)mdl * dd;
        const double *tab = p.pow + (size_t)mdl * p.powCap * dd;
        const int size = p.meta[2 * mdl], steady = p.meta[2 * mdl + 1];
        const int steadyN = steady ? size : -1;                          // RpowSteadyN :460-473
        const double t0 = p.tBottom[br], t1 = p.tTop[br];
        const int a = p.typeBottom[br], b = p.typeTop[br];
        const double L = sub_rn(t1, t0);                                 // :124
        const double muL = mul_rn(p.mu[mdl], L);                         // :130
        const double *E = expm_QL(d, Q, L, buf, rc, lane);                   // :132-134
        const double Pba = E[a * d + b];
        __syncwarp();

        int st = MTT_ST_OK, nVirt = -1, nChg = 0;
        double lp = -INFINITY;
        if (Pba == 0.0)                                                  // :137-138
            st = MTT_ST_NO_VALID_PATH;
        else if (Pba > 1.0 || Pba < 0.0)                                 // :141-144
            st = MTT_ST_PBA_RUBBISH;

        // ---- drawEventCount (:74-105)
        if (st == MTT_ST_OK) {
            const int forced = p.nVirtForced ? p.nVirtForced[br] : -1;
            if (forced >= 0) {
                nVirt = forced;
            } else if (pos >= nU) {
                st = MTT_ST_NEED_MORE_UNIFORMS;
            } else {
                const double uu = u[pos++];                              // :79
                double Plow = 0.0;
                double acc = sub_rn(-muL, log_ni(Pba));                     // :81
                const double log_muL = log_ni(muL);                         // :82
                int n = 0;
                while (true) {
                    if (n >= size && !steady) {
                        st = MTT_ST_NEED_MORE_POWERS;
                        break;
                    }
                    const double Rn = tab[(size_t)min(n, size - 1) * dd + a * d + b];
                    Plow = add_rn(Plow, exp_ni(add_rn(log_ni(Rn), acc)));      // :86
                    if (Plow > uu) {                                     // :88-89
                        nVirt = n;
                        break;
                    }
                    n += 1;
                    acc = add_rn(acc, sub_rn(log_muL, log_ni((double)n)));  // :92
                    if (!(steadyN < 0 || n < steadyN)) {                 // :94
                        st = MTT_ST_NEEDS_POISSON_TAIL;                  // :96-102 -> host
                        nVirt = n;                                       // thresh
                        break;
                    }
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
        if (st == MTT_ST_OK && nVirt > U_TCAP) {
            if (nVirt > p.scratchCap) {
                st = MTT_ST_NEED_MORE_POWERS;   // host grows the scratch along with the table
            } else {
                const size_t w = (size_t)blockIdx.x * U_WARPS + warp;
                times = p.scratchT + w * p.scratchCap;
                types = p.scratchTy + w * p.scratchCap;
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
            int prev = a;
            for (int i = 1; i <= nVirt; i++) {
                const double *Rhi = tab + (size_t)min(nVirt - i + 1, size - 1) * dd;
                const double *Rlo = tab + (size_t)min(nVirt - i, size - 1) * dd;
                double u2 = mul_rn(u[pos++], Rhi[prev * d + b]);         // :165-166
                double prod = 0.0;
                if (lane < d)
                    prod = mul_rn(R[prev * d + lane], Rlo[lane * d + b]);
                int c = 0;
                bool fell = true;
                for (c = 0; c < d; c++) {
                    u2 = sub_rn(u2, __shfl_sync(FULL_MASK, prod, c));    // :170-171
                    if (u2 < 0.0) {
                        fell = false;
                        break;
                    }
                }
                if (fell) {                                              // :179-190
                    st = MTT_ST_FB_FELL_THROUGH;
                    break;
                }
                if (lane == 0)
                    types[i - 1] = (uint8_t)c;
                prev = c;
            }
            __syncwarp();
        }

        if (st == MTT_ST_OK) {
            // ---- real changes + path probability (:196-221)
            lp = 0.0;
            int prevType = a;
            double prevTime = t0;
            for (int i = 0; i < nVirt; i++) {
                const int ty = types[i];
                if (ty != prevType) {
                    const double ti = times[i];
                    if (lane == 0 && nChg < p.outCap) {
                        p.outChgTime[br * p.outCap + nChg] = ti;
                        p.outChgType[br * p.outCap + nChg] = (uint8_t)ty;
                    }
                    nChg++;
                    lp = add_rn(lp, add_rn(mul_rn(Q[prevType * d + prevType], sub_rn(ti, prevTime)),
                                           log_ni(Q[prevType * d + ty])));
                    prevType = ty;
                    prevTime = ti;
                }
            }
            lp = add_rn(lp, mul_rn(Q[prevType * d + prevType], sub_rn(t1, prevTime)));
            lp = sub_rn(lp, log_ni(Pba));
        }
        __syncwarp();
        if (lane == 0) {
            p.status[br] = st;
            p.nChg[br] = (st == MTT_ST_OK) ? nChg : 0;
            p.logProb[br] = (st == MTT_ST_OK) ? lp : -INFINITY;
            p.nVirt[br] = nVirt;
        }
        if (st != MTT_ST_OK)
            aborted = true;
    }
    if (lane == 0)
        p.uUsed[job] = pos;
}

// ------------------------------------------------------------------ host side

int ensure_powers(mtt_ctx *ctx, int K)
{
    const int nM = ctx->unifModels, d = ctx->unifD, dd = d * d;
    if (nM == 0)
        return mtt_fail(ctx, MTT_ERR_STATE, "mtt_unif_set_models has not been called");
    bool need = false;
    for (int i = 0; i < nM; i++)
        if (!ctx->unifSteady[i] && ctx->unifPowSize[i] <= K)
            need = true;
    if (!need)
        return MTT_OK;
    cudaStream_t s = ctx->stream;
    if (K + 1 > ctx->unifPowCap) {
        // grow the table, keeping the powers already computed
        int newCap = std::max(K + 1 + 16, 2 * ctx->unifPowCap);
        DevBuf &b = ctx->bufs[BUF_U_POW];
        double *old = (double *)b.p;
        const int oldCap = ctx->unifPowCap;
        double *nw = nullptr;
        MTT_CUDA(ctx, cudaMalloc(&nw, (size_t)nM * newCap * dd * 8 + 256));
        if (old) {
            MTT_CUDA(ctx, cudaMemcpy2DAsync(nw, (size_t)newCap * dd * 8, old,
                                            (size_t)oldCap * dd * 8, (size_t)oldCap * dd * 8, nM,
                                            cudaMemcpyDeviceToDevice, s));
            MTT_CUDA(ctx, cudaStreamSynchronize(s));
            cudaFree(old);
        }
        b.p = nw;
        b.cap = (size_t)nM * newCap * dd * 8 + 256;
        ctx->unifPowCap = newCap;
    }
    int *meta = (int *)ctx->bufs[BUF_U_META].p;
    const int blocks = (nM + U_WARPS - 1) / U_WARPS;
    k3_rpow_extend<<<blocks, U_WARPS * 32, (size_t)U_WARPS * 3 * dd * 8, s>>>(
        nM, d, (const double *)ctx->bufs[BUF_U_R].p, (double *)ctx->bufs[BUF_U_POW].p,
        ctx->unifPowCap, meta, K);
    MTT_CHECK_LAUNCH(ctx);
    std::vector<int> h(2 * nM);
    MTT_CUDA(ctx, cudaMemcpyAsync(h.data(), meta, sizeof(int) * 2 * nM, cudaMemcpyDeviceToHost, s));
    MTT_CUDA(ctx, cudaStreamSynchronize(s));
    for (int i = 0; i < nM; i++) {
        ctx->unifPowSize[i] = h[2 * i];
        ctx->unifSteady[i] = h[2 * i + 1];
    }
    return MTT_OK;
}

size_t expm_smem(int d) { return (size_t)U_WARPS * 8 * d * d * 8; }

}  // namespace

extern "C" {

int mtt_unif_set_models(mtt_ctx *ctx, int nModels, int d, const double *m, int sym)
{
    if (!ctx)
        return MTT_ERR_INVALID_ARG;
    if (nModels < 1 || !m)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "mtt_unif_set_models: bad arguments");
    if (d < 2 || d > MTT_MAX_DEMES)
        return mtt_fail(ctx, MTT_ERR_TOO_LARGE, "uniformization: d=%d outside [2,%d]", d,
                        MTT_MAX_DEMES);
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t dd = (size_t)d * d;
    double *dM = (double *)mtt_buf(ctx, BUF_U_IN0, nModels * dd * 8);
    double *dQ = (double *)mtt_buf(ctx, BUF_U_Q, nModels * dd * 8);
    double *dR = (double *)mtt_buf(ctx, BUF_U_R, nModels * dd * 8);
    double *dMu = (double *)mtt_buf(ctx, BUF_U_MU, nModels * 8);
    int *meta = (int *)mtt_buf(ctx, BUF_U_META, nModels * 2 * sizeof(int));
    if (!dM || !dQ || !dR || !dMu || !meta)
        return MTT_ERR_CUDA;
    MTT_CUDA(ctx, cudaMemcpyAsync(dM, m, nModels * dd * 8, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemsetAsync(meta, 0, nModels * 2 * sizeof(int), s));   // drop cached powers
    const int wpb = 4;
    k3_update_matrices<<<(nModels + wpb - 1) / wpb, wpb * 32, 0, s>>>(nModels, d, dM, sym, dQ, dR,
                                                                      dMu);
    MTT_CHECK_LAUNCH(ctx);
    if (ctx->unifModels != nModels || ctx->unifD != d) {
        // table geometry changed: release it
        DevBuf &b = ctx->bufs[BUF_U_POW];
        if (b.p) {
            MTT_CUDA(ctx, cudaStreamSynchronize(s));
            cudaFree(b.p);
            b.p = nullptr;
            b.cap = 0;
        }
        ctx->unifPowCap = 0;
    }
    ctx->unifModels = nModels;
    ctx->unifD = d;
    ctx->unifSym = sym;
    ctx->unifPowSize.assign(nModels, 0);
    ctx->unifSteady.assign(nModels, 0);
    ctx->unifMu.resize(nModels);
    MTT_CUDA(ctx, cudaMemcpyAsync(ctx->unifMu.data(), dMu, nModels * 8, cudaMemcpyDeviceToHost, s));
    MTT_CUDA(ctx, cudaStreamSynchronize(s));
    return MTT_OK;
}

int mtt_unif_get_matrices(mtt_ctx *ctx, int model, double *Q, double *R, double *mu)
{
    if (!ctx)
        return MTT_ERR_INVALID_ARG;
    if (model < 0 || model >= ctx->unifModels)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "model %d out of range", model);
    const size_t dd = (size_t)ctx->unifD * ctx->unifD;
    cudaStream_t s = ctx->stream;
    if (Q)
        MTT_CUDA(ctx, cudaMemcpyAsync(Q, (double *)ctx->bufs[BUF_U_Q].p + model * dd, dd * 8,
                                      cudaMemcpyDeviceToHost, s));
    if (R)
        MTT_CUDA(ctx, cudaMemcpyAsync(R, (double *)ctx->bufs[BUF_U_R].p + model * dd, dd * 8,
                                      cudaMemcpyDeviceToHost, s));
    MTT_CUDA(ctx, cudaStreamSynchronize(s));
    if (mu)
        *mu = ctx->unifMu[model];
    return MTT_OK;
}

int mtt_unif_rpow(mtt_ctx *ctx, int model, int kmax, double *out, int *steadyN)
{
    if (!ctx)
        return MTT_ERR_INVALID_ARG;
    if (model < 0 || model >= ctx->unifModels || kmax < 0 || !out)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "mtt_unif_rpow: bad arguments");
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = ensure_powers(ctx, kmax);
    if (rc != MTT_OK)
        return rc;
    const size_t dd = (size_t)ctx->unifD * ctx->unifD;
    const int size = ctx->unifPowSize[model];
    const int have = std::min(kmax + 1, size);
    const double *tab = (const double *)ctx->bufs[BUF_U_POW].p + (size_t)model * ctx->unifPowCap * dd;
    MTT_CUDA(ctx, cudaMemcpyAsync(out, tab, have * dd * 8, cudaMemcpyDeviceToHost, ctx->stream));
    MTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = have; k <= kmax; k++)          // steady: later powers repeat the last (:422-426)
        memcpy(out + k * dd, out + (size_t)(size - 1) * dd, dd * 8);
    if (steadyN)
        *steadyN = ctx->unifSteady[model] ? size : -1;
    return MTT_OK;
}

int mtt_unif_expm_batch(mtt_ctx *ctx, int64_t n, const int32_t *model, const double *L,
                        double *out)
{
    if (!ctx)
        return MTT_ERR_INVALID_ARG;
    if (ctx->unifModels == 0)
        return mtt_fail(ctx, MTT_ERR_STATE, "mtt_unif_set_models has not been called");
    if (n < 0 || !model || !L || !out)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "mtt_unif_expm_batch: NULL argument");
    if (n == 0)
        return MTT_OK;
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int d = ctx->unifD;
    const size_t dd = (size_t)d * d;
    int *dModel = (int *)mtt_buf(ctx, BUF_U_IN0, n * 4);
    double *dL = (double *)mtt_buf(ctx, BUF_U_IN1, n * 8);
    double *dOut = (double *)mtt_buf(ctx, BUF_U_OUT0, n * dd * 8);
    if (!dModel || !dL || !dOut)
        return MTT_ERR_CUDA;
    MTT_CUDA(ctx, cudaMemcpyAsync(dModel, model, n * 4, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dL, L, n * 8, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaFuncSetAttribute(k3_expm_batch, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)expm_smem(d)));
    k3_expm_batch<<<(unsigned)((n + U_WARPS - 1) / U_WARPS), U_WARPS * 32, expm_smem(d), s>>>(
        n, d, (const double *)ctx->bufs[BUF_U_Q].p, dModel, dL, dOut);
    MTT_CHECK_LAUNCH(ctx);
    MTT_CUDA(ctx, cudaMemcpyAsync(out, dOut, n * dd * 8, cudaMemcpyDeviceToHost, s));
    MTT_CUDA(ctx, cudaStreamSynchronize(s));
    return MTT_OK;
}

int mtt_unif_branch_logp_batch(mtt_ctx *ctx, int64_t n, const int32_t *model,
                               const double *tBottom, const double *tTop,
                               const uint8_t *typeBottom, const uint8_t *typeTop,
                               const int64_t *chgOff, const double *chgTime,
                               const uint8_t *chgType, double *logProb, int32_t *status)
{
    if (!ctx)
        return MTT_ERR_INVALID_ARG;
    if (ctx->unifModels == 0)
        return mtt_fail(ctx, MTT_ERR_STATE, "mtt_unif_set_models has not been called");
    if (n < 0 || !model || !tBottom || !tTop || !typeBottom || !typeTop || !chgOff || !logProb ||
        !status)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "mtt_unif_branch_logp_batch: NULL argument");
    if (n == 0)
        return MTT_OK;
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int d = ctx->unifD;
    const int64_t nChg = chgOff[n];
    int *dModel = (int *)mtt_buf(ctx, BUF_U_IN0, n * 4);
    double *dT0 = (double *)mtt_buf(ctx, BUF_U_IN1, n * 8);
    double *dT1 = (double *)mtt_buf(ctx, BUF_U_IN2, n * 8);
    uint8_t *dA = (uint8_t *)mtt_buf(ctx, BUF_U_IN3, n);
    uint8_t *dB = (uint8_t *)mtt_buf(ctx, BUF_U_IN4, n);
    long long *dOff = (long long *)mtt_buf(ctx, BUF_U_IN5, (n + 1) * 8);
    double *dCT = (double *)mtt_buf(ctx, BUF_U_IN6, nChg * 8);
    uint8_t *dCTy = (uint8_t *)mtt_buf(ctx, BUF_U_IN7, nChg);
    double *dLP = (double *)mtt_buf(ctx, BUF_U_OUT0, n * 8);
    int *dSt = (int *)mtt_buf(ctx, BUF_U_OUT1, n * 4);
    if (!dModel || !dT0 || !dT1 || !dA || !dB || !dOff || !dCT || !dCTy || !dLP || !dSt)
        return MTT_ERR_CUDA;
    MTT_CUDA(ctx, cudaMemcpyAsync(dModel, model, n * 4, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dT0, tBottom, n * 8, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dT1, tTop, n * 8, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dA, typeBottom, n, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dB, typeTop, n, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dOff, chgOff, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    if (nChg > 0) {
        MTT_CUDA(ctx, cudaMemcpyAsync(dCT, chgTime, nChg * 8, cudaMemcpyHostToDevice, s));
        MTT_CUDA(ctx, cudaMemcpyAsync(dCTy, chgType, nChg, cudaMemcpyHostToDevice, s));
    }
    MTT_CUDA(ctx, cudaFuncSetAttribute(k3_branch_logp, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)expm_smem(d)));
    k3_branch_logp<<<(unsigned)((n + U_WARPS - 1) / U_WARPS), U_WARPS * 32, expm_smem(d), s>>>(
        n, d, (const double *)ctx->bufs[BUF_U_Q].p, dModel, dT0, dT1, dA, dB, dOff, dCT, dCTy, dLP,
        dSt);
    MTT_CHECK_LAUNCH(ctx);
    MTT_CUDA(ctx, cudaMemcpyAsync(logProb, dLP, n * 8, cudaMemcpyDeviceToHost, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(status, dSt, n * 4, cudaMemcpyDeviceToHost, s));
    MTT_CUDA(ctx, cudaStreamSynchronize(s));
    return MTT_OK;
}

int mtt_unif_retype_batch(mtt_ctx *ctx, int64_t nJobs, const int64_t *brOff,
                          const int32_t *model, const double *tBottom, const double *tTop,
                          const uint8_t *typeBottom, const uint8_t *typeTop,
                          const int32_t *nVirtForced, const double *uniforms,
                          const int64_t *uOff, int outCap, int32_t *nChg, double *outChgTime,
                          uint8_t *outChgType, double *logProb, int32_t *status, int32_t *nVirt,
                          int64_t *uUsed)
{
    if (!ctx)
        return MTT_ERR_INVALID_ARG;
    if (ctx->unifModels == 0)
        return mtt_fail(ctx, MTT_ERR_STATE, "mtt_unif_set_models has not been called");
    if (nJobs < 0 || !brOff || !model || !tBottom || !tTop || !typeBottom || !typeTop ||
        !uniforms || !uOff || outCap < 0 || !nChg || !logProb || !status || !nVirt || !uUsed ||
        (outCap > 0 && (!outChgTime || !outChgType)))
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "mtt_unif_retype_batch: NULL argument");
    if (nJobs == 0)
        return MTT_OK;
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int d = ctx->unifD;
    const size_t dd = (size_t)d * d;
    const int64_t nBr = brOff[nJobs], nU = uOff[nJobs];

    // powers needed: a generous Poisson bound on the number of virtual events per branch
    double maxMuL = 0.0;
    for (int64_t i = 0; i < nBr; i++) {
        if (model[i] < 0 || model[i] >= ctx->unifModels)
            return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "branch %lld: model out of range",
                            (long long)i);
        double muL = ctx->unifMu[model[i]] * (tTop[i] - tBottom[i]);
        if (muL > maxMuL)
            maxMuL = muL;
        if (nVirtForced && nVirtForced[i] > maxMuL)
            maxMuL = nVirtForced[i];
    }
    if (!(maxMuL < 1e7))
        return mtt_fail(ctx, MTT_ERR_TOO_LARGE, "mu*L = %g is too large", maxMuL);
    int K = (int)(maxMuL + 10.0 * sqrt(maxMuL) + 32.0);

    int *dModel = (int *)mtt_buf(ctx, BUF_U_IN0, nBr * 4);
    double *dT0 = (double *)mtt_buf(ctx, BUF_U_IN1, nBr * 8);
    double *dT1 = (double *)mtt_buf(ctx, BUF_U_IN2, nBr * 8);
    uint8_t *dA = (uint8_t *)mtt_buf(ctx, BUF_U_IN3, nBr);
    uint8_t *dB = (uint8_t *)mtt_buf(ctx, BUF_U_IN4, nBr);
    long long *dBrOff = (long long *)mtt_buf(ctx, BUF_U_IN5, (nJobs + 1) * 8);
    double *dU = (double *)mtt_buf(ctx, BUF_U_IN6, nU * 8);
    long long *dUOff = (long long *)mtt_buf(ctx, BUF_U_IN7, (nJobs + 1) * 8);
    int *dForced = nVirtForced ? (int *)mtt_buf(ctx, BUF_U_IN8, nBr * 4) : nullptr;
    int *dNChg = (int *)mtt_buf(ctx, BUF_U_OUT0, nBr * 4);
    double *dOCT = (double *)mtt_buf(ctx, BUF_U_OUT1, (size_t)nBr * outCap * 8);
    uint8_t *dOCTy = (uint8_t *)mtt_buf(ctx, BUF_U_OUT2, (size_t)nBr * outCap);
    double *dLP = (double *)mtt_buf(ctx, BUF_U_OUT3, nBr * 8);
    int *dSt = (int *)mtt_buf(ctx, BUF_U_OUT4, nBr * 4 * 2);     // status | nVirt
    long long *dUsed = (long long *)mtt_buf(ctx, BUF_U_OUT5, nJobs * 8);
    if (!dModel || !dT0 || !dT1 || !dA || !dB || !dBrOff || !dU || !dUOff || !dNChg || !dOCT ||
        !dOCTy || !dLP || !dSt || !dUsed || (nVirtForced && !dForced))
        return MTT_ERR_CUDA;
    int *dNVirt = dSt + nBr;
    MTT_CUDA(ctx, cudaMemcpyAsync(dModel, model, nBr * 4, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dT0, tBottom, nBr * 8, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dT1, tTop, nBr * 8, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dA, typeBottom, nBr, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dB, typeTop, nBr, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dBrOff, brOff, (nJobs + 1) * 8, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dU, uniforms, nU * 8, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dUOff, uOff, (nJobs + 1) * 8, cudaMemcpyHostToDevice, s));
    if (nVirtForced)
        MTT_CUDA(ctx, cudaMemcpyAsync(dForced, nVirtForced, nBr * 4, cudaMemcpyHostToDevice, s));

    const unsigned blocks = (unsigned)((nJobs + U_WARPS - 1) / U_WARPS);
    const size_t smem = (size_t)U_WARPS * (8 * dd + U_TCAP + U_TCAP / 8) * 8;
    MTT_CUDA(ctx, cudaFuncSetAttribute(k3_retype, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    std::vector<int32_t> hSt(nBr);
    for (int attempt = 0; attempt < 12; attempt++) {
        int rc = ensure_powers(ctx, K);
        if (rc != MTT_OK)
            return rc;
        int scratchCap = 0;
        double *scrT = nullptr;
        uint8_t *scrTy = nullptr;
        if (K > U_TCAP) {
            scratchCap = K + 1;
            scrT = (double *)mtt_buf(ctx, BUF_U_OUT6, (size_t)blocks * U_WARPS * scratchCap * 9);
            if (!scrT)
                return MTT_ERR_CUDA;
            scrTy = (uint8_t *)(scrT + (size_t)blocks * U_WARPS * scratchCap);
        }
        RetypeParams p;
        p.nJobs = nJobs; p.d = d; p.powCap = ctx->unifPowCap; p.outCap = outCap;
        p.Q = (const double *)ctx->bufs[BUF_U_Q].p; p.R = (const double *)ctx->bufs[BUF_U_R].p;
        p.mu = (const double *)ctx->bufs[BUF_U_MU].p; p.pow = (const double *)ctx->bufs[BUF_U_POW].p;
        p.meta = (const int *)ctx->bufs[BUF_U_META].p;
        p.brOff = dBrOff; p.model = dModel; p.tBottom = dT0; p.tTop = dT1; p.typeBottom = dA;
        p.typeTop = dB; p.nVirtForced = dForced; p.uniforms = dU; p.uOff = dUOff;
        p.nChg = dNChg; p.outChgTime = dOCT; p.outChgType = dOCTy; p.logProb = dLP;
        p.status = dSt; p.nVirt = dNVirt; p.uUsed = dUsed;
        p.scratchT = scrT; p.scratchTy = scrTy; p.scratchCap = scratchCap;
        k3_retype<<<blocks, U_WARPS * 32, smem, s>>>(p);
        MTT_CHECK_LAUNCH(ctx);
        MTT_CUDA(ctx, cudaMemcpyAsync(hSt.data(), dSt, nBr * 4, cudaMemcpyDeviceToHost, s));
        MTT_CUDA(ctx, cudaStreamSynchronize(s));
        bool more = false;
        for (int64_t i = 0; i < nBr; i++)
            if (hSt[i] == MTT_ST_NEED_MORE_POWERS)
                more = true;
        if (!more)
            break;
        K = 2 * K + 64;      // the table was too short for some draw: extend and redo the batch
        if (attempt == 11)
            return mtt_fail(ctx, MTT_ERR_TOO_LARGE, "R^k table would exceed %d powers", K);
    }
    memcpy(status, hSt.data(), nBr * 4);
    MTT_CUDA(ctx, cudaMemcpyAsync(nChg, dNChg, nBr * 4, cudaMemcpyDeviceToHost, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(nVirt, dNVirt, nBr * 4, cudaMemcpyDeviceToHost, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(logProb, dLP, nBr * 8, cudaMemcpyDeviceToHost, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(uUsed, dUsed, nJobs * 8, cudaMemcpyDeviceToHost, s));
    if (outCap > 0) {
        MTT_CUDA(ctx, cudaMemcpyAsync(outChgTime, dOCT, (size_t)nBr * outCap * 8,
                                      cudaMemcpyDeviceToHost, s));
        MTT_CUDA(ctx, cudaMemcpyAsync(outChgType, dOCTy, (size_t)nBr * outCap,
                                      cudaMemcpyDeviceToHost, s));
    }
    MTT_CUDA(ctx, cudaStreamSynchronize(s));
    return MTT_OK;
}

}  // extern "C"
