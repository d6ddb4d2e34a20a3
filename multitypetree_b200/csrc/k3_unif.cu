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
#include "k3_device.cuh"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace {

// ---- updateMatrices (SCMigrationModel.java:150-201), one warp per model ------------------
__global__ void k3_update_matrices(int nModels, int d, const double *m, int sym, double *Q,
                                   double *R, double *mu, double *logQ)
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
    // log Q[i,j], the term a type change i -> j adds to a path probability (:211-221, :249-258):
    // tabulated once per model with the device log() the kernels would call per change
    double *lq = logQ + (size_t)model * d * d;
    #pragma unroll 1
    for (int e = lane; e < d * d; e += 32)
        lq[e] = log_ni(q[e]);
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
        mm<false>(d, sPrev, sR, sNew, rc, lane);
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

// elementwise log of the valid part of every model's R^k table (drawEventCount reads
// log(R^n[a,b]) once per candidate n: tabulating it takes the log out of the sampler's loop)
__global__ void k3_log_table(int nModels, int dd, int powCap, const int *meta, const double *pow,
                             double *logPow)
{
    const int model = blockIdx.y;
    const long long n = (long long)meta[2 * model] * dd;
    const size_t base = (size_t)model * powCap * dd;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x)
        logPow[base + e] = log_ni(pow[base + e]);
}

__global__ void k3_log_n(double *logN)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < U_LOGN)
        logN[n] = log_ni((double)n);
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
    const double *F = expm_QL<true>(d, Q + (size_t)model[item] * dd, L[item], buf, rc, lane);
    #pragma unroll 1
    for (int e = lane; e < dd; e += 32)
        out[item * dd + e] = F[e];
}

// expm(Q L)[a,b] of every branch of a retype batch, one warp per branch: the three matrix
// exponentials of a NodeRetype job do not depend on its random numbers, so they run side by side
// here instead of one after the other inside the job's (sequential) sampling warp
__global__ void k3_pba(long long n, int d, const double *Qall, const int *model,
                       const double *tBottom, const double *tTop, const uint8_t *typeBottom,
                       const uint8_t *typeTop, double *pba)
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
    const double L = sub_rn(tTop[br], tBottom[br]);
    const double *E = expm_QL<true>(d, Qall + (size_t)model[br] * dd, L, buf, rc, lane);
    if (lane == 0)
        pba[br] = E[typeBottom[br] * d + typeTop[br]];
}

// ---- getBranchTypeProb (UniformizationRetypeOperator.java:233-275), one warp per branch ----
__global__ void k3_branch_logp(long long n, int d, const double *Qall, const double *logQall,
                               const int *model,
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
    const double *Q = Qall + (size_t)model[br] * dd, *lQ = logQall + (size_t)model[br] * dd;
    const double t0 = tBottom[br], t1 = tTop[br];
    const int a = typeBottom[br], b = typeTop[br];
    double lp = 0.0, lastTime = t0;
    int lastCol = a;
    for (long long k = chgOff[br]; k < chgOff[br + 1]; k++) {           // :249-258
        const double thisTime = chgTime[k];
        const int thisCol = chgType[k];
        lp = add_rn(lp, add_rn(mul_rn(sub_rn(thisTime, lastTime), Q[lastCol * d + lastCol]),
                               lQ[lastCol * d + thisCol]));
        lastTime = thisTime;
        lastCol = thisCol;
    }
    lp = add_rn(lp, mul_rn(sub_rn(t1, lastTime), Q[lastCol * d + lastCol]));   // :259
    const double *E = expm_QL<true>(d, Q, sub_rn(t1, t0), buf, rc, lane);    // :263-264
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
    const double *Q, *R, *mu, *pow, *logPow, *logN, *logQ;
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
    const double *pba;       // expm(Q L)[a,b] per branch (k3_pba), or NULL
    double *scratchT;        // [warps][scratchCap] spill for nVirt > U_TCAP
    uint8_t *scratchTy;
    int scratchCap;
    int tCap;                // virtual events kept in shared memory per warp (<= U_TCAP)
};

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
    const size_t perWarp = (size_t)8 * dd + p.tCap + p.tCap / 8;
    double *buf = sm + (size_t)warp * perWarp;
    double *sTimes = buf + 8 * dd;
    uint8_t *sTypes = reinterpret_cast<uint8_t *>(sTimes + p.tCap);

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
        const double *lQ = p.logQ + (size_t)mdl * dd;
        const double t0 = p.tBottom[br], t1 = p.tTop[br];
        const int a = p.typeBottom[br], b = p.typeTop[br];
        const size_t w = (size_t)blockIdx.x * U_WARPS + warp;
        const BranchSample bs = sample_branch<false, true>(
            d, Q, R, p.pow + (size_t)mdl * p.powCap * dd, p.logPow + (size_t)mdl * p.powCap * dd,
            p.logN, p.meta[2 * mdl], p.meta[2 * mdl + 1],
            p.mu[mdl], t0, t1, a, b, p.nVirtForced ? p.nVirtForced[br] : -1, u, nU, pos, buf, rc,
            sTimes, sTypes, p.scratchT + w * p.scratchCap, p.scratchTy + w * p.scratchCap,
            p.scratchCap, lane, p.tCap, p.pba ? p.pba + br : nullptr);
        int st = bs.st, nChg = 0;
        const int nVirt = bs.nVirt;
        const double Pba = bs.Pba;
        const double *times = bs.times;
        const uint8_t *types = bs.types;
        double lp = -INFINITY;

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
                                           lQ[prevType * d + ty]));
                    prevType = ty;
                    prevTime = ti;
                }
            }
            lp = add_rn(lp, mul_rn(Q[prevType * d + prevType], sub_rn(t1, prevTime)));
            lp = sub_rn(lp, isnan(bs.logPba) ? log_ni(Pba) : bs.logPba);
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
    double *logPow = (double *)mtt_buf(ctx, BUF_U_LOGPOW, (size_t)nM * ctx->unifPowCap * dd * 8);
    if (!logPow)
        return MTT_ERR_CUDA;
    k3_log_table<<<dim3(32, (unsigned)nM), 256, 0, s>>>(nM, dd, ctx->unifPowCap, meta,
                                                        (const double *)ctx->bufs[BUF_U_POW].p,
                                                        logPow);
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

int mtt_unif_ensure_powers(mtt_ctx *ctx, int K)
{
    return ensure_powers(ctx, K);
}

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
    double *dLogQ = (double *)mtt_buf(ctx, BUF_U_LOGQ, nModels * dd * 8);
    if (!dLogQ)
        return MTT_ERR_CUDA;
    k3_update_matrices<<<(nModels + wpb - 1) / wpb, wpb * 32, 0, s>>>(nModels, d, dM, sym, dQ, dR,
                                                                      dMu, dLogQ);
    MTT_CHECK_LAUNCH(ctx);
    double *logN = (double *)mtt_buf(ctx, BUF_U_LOGN, U_LOGN * 8);
    if (!logN)
        return MTT_ERR_CUDA;
    k3_log_n<<<U_LOGN / 256, 256, 0, s>>>(logN);
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
    MTT_CUDA(ctx, mtt_fn_smem(ctx, k3_expm_batch, expm_smem(d)));
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
    int *dModel;
    double *dT0, *dT1, *dCT, *dLP;
    uint8_t *dA, *dB, *dCTy;
    long long *dOff;
    int *dSt;
    Xfer x(ctx);
    x.in(model, (size_t)n, BUF_U_IN0, &dModel);
    x.in(tBottom, (size_t)n, BUF_U_IN1, &dT0);
    x.in(tTop, (size_t)n, BUF_U_IN2, &dT1);
    x.in(typeBottom, (size_t)n, BUF_U_IN3, &dA);
    x.in(typeTop, (size_t)n, BUF_U_IN4, &dB);
    x.in((const long long *)chgOff, (size_t)n + 1, BUF_U_IN5, &dOff);
    x.in(chgTime, (size_t)nChg, BUF_U_IN6, &dCT);
    x.in(chgType, (size_t)nChg, BUF_U_IN7, &dCTy);
    x.out(logProb, (size_t)n, BUF_U_OUT0, &dLP);
    x.out(status, (size_t)n, BUF_U_OUT1, &dSt);
    x.direct_out();
    int rc = x.prepare();
    if (rc == MTT_OK)
        rc = x.upload();
    if (rc != MTT_OK)
        return rc;
    MTT_CUDA(ctx, mtt_fn_smem(ctx, k3_branch_logp, expm_smem(d)));
    k3_branch_logp<<<(unsigned)((n + U_WARPS - 1) / U_WARPS), U_WARPS * 32, expm_smem(d), s>>>(
        n, d, (const double *)ctx->bufs[BUF_U_Q].p, (const double *)ctx->bufs[BUF_U_LOGQ].p, dModel,
        dT0, dT1, dA, dB, dOff, dCT, dCTy, dLP,
        dSt);
    MTT_CHECK_LAUNCH(ctx);
    return x.download();
}

}  // extern "C"

namespace {

// the branches whose CURRENT typing is evaluated in the same call (mtt_unif_node_retype_batch)
struct OldTyping {
    int64_t n;
    const int32_t *model;
    const double *tBottom, *tTop;
    const uint8_t *typeBottom, *typeTop;
    const int64_t *chgOff;
    const double *chgTime;
    const uint8_t *chgType;
    double *logProb;
    int32_t *status;
};

constexpr int RETYPE_REDO = 1;        // (internal) the R^k table was too short

int retype_chunked(mtt_ctx *ctx, int64_t nJobs, const int64_t *brOff, const int32_t *model,
                   const double *tBottom, const double *tTop, const uint8_t *typeBottom,
                   const uint8_t *typeTop, const int32_t *nVirtForced, const double *uniforms,
                   const int64_t *uOff, int outCap, int32_t *nChg, double *outChgTime,
                   uint8_t *outChgType, double *logProb, int32_t *status, int32_t *nVirt,
                   int64_t *uUsed, int K)
{
    constexpr int NC = 8;
    cudaStream_t s = ctx->stream;
    const int d = ctx->unifD;
    const size_t dd = (size_t)d * d;
    const int64_t nBr = brOff[nJobs], nU = uOff[nJobs];
    if (!ctx->copyStream)
        MTT_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
    if (!ctx->copyStream2)
        MTT_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copyStream2, cudaStreamNonBlocking));
    while ((int)ctx->events.size() < 2 * NC + 1) {
        cudaEvent_t ev;
        MTT_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->events.push_back(ev);
    }
    // the per-branch and per-job inputs are small next to the uniforms: one upload up front
    int *dModel, *dForced = nullptr;
    double *dT0, *dT1;
    uint8_t *dA, *dB;
    long long *dBrOff, *dUOff;
    Xfer x(ctx, 0);                     // (never staged: these arrays are megabytes)
    x.in(model, (size_t)nBr, BUF_U_IN0, &dModel);
    x.in(tBottom, (size_t)nBr, BUF_U_IN1, &dT0);
    x.in(tTop, (size_t)nBr, BUF_U_IN2, &dT1);
    x.in(typeBottom, (size_t)nBr, BUF_U_IN3, &dA);
    x.in(typeTop, (size_t)nBr, BUF_U_IN4, &dB);
    x.in((const long long *)brOff, (size_t)nJobs + 1, BUF_U_IN5, &dBrOff);
    x.in((const long long *)uOff, (size_t)nJobs + 1, BUF_U_IN7, &dUOff);
    if (nVirtForced)
        x.in(nVirtForced, (size_t)nBr, BUF_U_IN8, &dForced);
    int rc = x.prepare();
    if (rc == MTT_OK)
        rc = x.upload();
    if (rc != MTT_OK)
        return rc;
    double *dU = (double *)mtt_buf(ctx, BUF_U_IN6, (size_t)nU * 8);
    int *dNChg = (int *)mtt_buf(ctx, BUF_U_OUT0, (size_t)nBr * 4);
    double *dOCT = (double *)mtt_buf(ctx, BUF_U_OUT1, (size_t)nBr * outCap * 8);
    uint8_t *dOCTy = (uint8_t *)mtt_buf(ctx, BUF_U_OUT2, (size_t)nBr * outCap);
    double *dLP = (double *)mtt_buf(ctx, BUF_U_OUT3, (size_t)nBr * 8);
    int *dSt = (int *)mtt_buf(ctx, BUF_U_OUT4, (size_t)nBr * 4);
    int *dNVirt = (int *)mtt_buf(ctx, BUF_U_OUT7, (size_t)nBr * 4);
    long long *dUsed = (long long *)mtt_buf(ctx, BUF_U_OUT5, (size_t)nJobs * 8);
    double *dPba = (double *)mtt_buf(ctx, BUF_U_PBA, (size_t)nBr * 8);
    if (!dU || !dNChg || !dOCT || !dOCTy || !dLP || !dSt || !dNVirt || !dUsed || !dPba)
        return MTT_ERR_CUDA;
    if ((rc = ensure_powers(ctx, K)) != MTT_OK)
        return rc;
    int tCap = 32;
    while (tCap < K && tCap < U_TCAP)
        tCap <<= 1;
    const size_t smem = (size_t)U_WARPS * (8 * dd + tCap + tCap / 8) * 8;
    MTT_CUDA(ctx, mtt_fn_smem(ctx, k3_retype, smem));
    MTT_CUDA(ctx, mtt_fn_smem(ctx, k3_pba, expm_smem(d)));
    const int64_t perChunk = (nJobs + NC - 1) / NC;
    const unsigned maxBlocks = (unsigned)((perChunk + U_WARPS - 1) / U_WARPS);
    int scratchCap = 0;
    double *scrT = nullptr;
    uint8_t *scrTy = nullptr;
    if (K > U_TCAP) {
        scratchCap = K + 1;
        scrT = (double *)mtt_buf(ctx, BUF_U_OUT6, (size_t)maxBlocks * U_WARPS * scratchCap * 9);
        if (!scrT)
            return MTT_ERR_CUDA;
        scrTy = (uint8_t *)(scrT + (size_t)maxBlocks * U_WARPS * scratchCap);
    }
    // everything above runs on the main stream: the copy streams start after it
    MTT_CUDA(ctx, cudaEventRecord(ctx->events[2 * NC], s));
    MTT_CUDA(ctx, cudaStreamWaitEvent(ctx->copyStream, ctx->events[2 * NC], 0));
    MTT_CUDA(ctx, cudaStreamWaitEvent(ctx->copyStream2, ctx->events[2 * NC], 0));
    for (int c = 0; c < NC; c++) {
        const int64_t j0 = std::min<int64_t>(nJobs, perChunk * c);
        const int64_t j1 = std::min<int64_t>(nJobs, perChunk * (c + 1));
        if (j1 <= j0)
            continue;
        const int64_t br0 = brOff[j0], br1 = brOff[j1], u0 = uOff[j0], u1 = uOff[j1];
        // up: this chunk's uniforms
        if (u1 > u0)
            MTT_CUDA(ctx, cudaMemcpyAsync(dU + u0, uniforms + u0, (size_t)(u1 - u0) * 8,
                                          cudaMemcpyHostToDevice, ctx->copyStream));
        MTT_CUDA(ctx, cudaEventRecord(ctx->events[c], ctx->copyStream));
        MTT_CUDA(ctx, cudaStreamWaitEvent(s, ctx->events[c], 0));
        // kernels: the branches' matrix exponentials, then the jobs
        if (br1 > br0) {
            k3_pba<<<(unsigned)((br1 - br0 + U_WARPS - 1) / U_WARPS), U_WARPS * 32, expm_smem(d),
                     s>>>(br1 - br0, d, (const double *)ctx->bufs[BUF_U_Q].p, dModel + br0,
                          dT0 + br0, dT1 + br0, dA + br0, dB + br0, dPba + br0);
            MTT_CHECK_LAUNCH(ctx);
        }
        RetypeParams p;
        p.nJobs = j1 - j0; p.d = d; p.powCap = ctx->unifPowCap; p.outCap = outCap;
        p.Q = (const double *)ctx->bufs[BUF_U_Q].p; p.R = (const double *)ctx->bufs[BUF_U_R].p;
        p.logQ = (const double *)ctx->bufs[BUF_U_LOGQ].p;
        p.mu = (const double *)ctx->bufs[BUF_U_MU].p; p.pow = (const double *)ctx->bufs[BUF_U_POW].p;
        p.logPow = (const double *)ctx->bufs[BUF_U_LOGPOW].p;
        p.logN = (const double *)ctx->bufs[BUF_U_LOGN].p;
        p.meta = (const int *)ctx->bufs[BUF_U_META].p;
        // (offsets stay absolute: the branch / uniform / output arrays are passed whole)
        p.brOff = dBrOff + j0; p.model = dModel; p.tBottom = dT0; p.tTop = dT1; p.typeBottom = dA;
        p.typeTop = dB; p.nVirtForced = dForced; p.uniforms = dU; p.uOff = dUOff + j0;
        p.nChg = dNChg; p.outChgTime = dOCT; p.outChgType = dOCTy; p.logProb = dLP;
        p.status = dSt; p.nVirt = dNVirt; p.uUsed = dUsed + j0;
        p.scratchT = scrT; p.scratchTy = scrTy; p.scratchCap = scratchCap; p.tCap = tCap;
        p.pba = dPba;
        k3_retype<<<(unsigned)((j1 - j0 + U_WARPS - 1) / U_WARPS), U_WARPS * 32, smem, s>>>(p);
        MTT_CHECK_LAUNCH(ctx);
        MTT_CUDA(ctx, cudaEventRecord(ctx->events[NC + c], s));
        // down: this chunk's results
        cudaStream_t dn = ctx->copyStream2;
        MTT_CUDA(ctx, cudaStreamWaitEvent(dn, ctx->events[NC + c], 0));
        const size_t nb = (size_t)(br1 - br0);
        MTT_CUDA(ctx, cudaMemcpyAsync(nChg + br0, dNChg + br0, nb * 4, cudaMemcpyDeviceToHost, dn));
        if (outCap > 0) {
            MTT_CUDA(ctx, cudaMemcpyAsync(outChgTime + br0 * outCap, dOCT + br0 * outCap,
                                          nb * outCap * 8, cudaMemcpyDeviceToHost, dn));
            MTT_CUDA(ctx, cudaMemcpyAsync(outChgType + br0 * outCap, dOCTy + br0 * outCap,
                                          nb * outCap, cudaMemcpyDeviceToHost, dn));
        }
        MTT_CUDA(ctx, cudaMemcpyAsync(logProb + br0, dLP + br0, nb * 8, cudaMemcpyDeviceToHost, dn));
        MTT_CUDA(ctx, cudaMemcpyAsync(status + br0, dSt + br0, nb * 4, cudaMemcpyDeviceToHost, dn));
        MTT_CUDA(ctx, cudaMemcpyAsync(nVirt + br0, dNVirt + br0, nb * 4, cudaMemcpyDeviceToHost, dn));
        MTT_CUDA(ctx, cudaMemcpyAsync(uUsed + j0, dUsed + j0, (size_t)(j1 - j0) * 8,
                                      cudaMemcpyDeviceToHost, dn));
    }
    MTT_CUDA(ctx, cudaStreamSynchronize(ctx->copyStream2));
    MTT_CUDA(ctx, cudaStreamSynchronize(s));
    for (int64_t i = 0; i < nBr; i++)
        if (status[i] == MTT_ST_NEED_MORE_POWERS)
            return RETYPE_REDO;
    return MTT_OK;
}

int retype_impl(mtt_ctx *ctx, int64_t nJobs, const int64_t *brOff,
                          const int32_t *model, const double *tBottom, const double *tTop,
                          const uint8_t *typeBottom, const uint8_t *typeTop,
                          const int32_t *nVirtForced, const double *uniforms,
                          const int64_t *uOff, int outCap, int32_t *nChg, double *outChgTime,
                          uint8_t *outChgType, double *logProb, int32_t *status, int32_t *nVirt,
                          int64_t *uUsed, const OldTyping *old)
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

    // Large batches (the throughput sweeps): the uniforms are most of the input and the change
    // lists most of the output, so the batch is cut into chunks of jobs whose upload, kernels
    // and download overlap on three streams.  A table that turns out too short (rare: K is a
    // generous bound) falls through to the one-piece path below, which extends it and redoes
    // the batch.
    if (!old && nJobs >= 16384) {
        const int rc = retype_chunked(ctx, nJobs, brOff, model, tBottom, tTop, typeBottom, typeTop,
                                      nVirtForced, uniforms, uOff, outCap, nChg, outChgTime,
                                      outChgType, logProb, status, nVirt, uUsed, K);
        if (rc != RETYPE_REDO)
            return rc;
        K = 2 * K + 64;
    }

    int *dModel, *dForced = nullptr, *dNChg, *dSt, *dNVirt;
    double *dT0, *dT1, *dU, *dOCT, *dLP;
    uint8_t *dA, *dB, *dOCTy;
    long long *dBrOff, *dUOff, *dUsed;
    Xfer x(ctx);
    x.in(model, (size_t)nBr, BUF_U_IN0, &dModel);
    x.in(tBottom, (size_t)nBr, BUF_U_IN1, &dT0);
    x.in(tTop, (size_t)nBr, BUF_U_IN2, &dT1);
    x.in(typeBottom, (size_t)nBr, BUF_U_IN3, &dA);
    x.in(typeTop, (size_t)nBr, BUF_U_IN4, &dB);
    x.in((const long long *)brOff, (size_t)nJobs + 1, BUF_U_IN5, &dBrOff);
    x.in(uniforms, (size_t)nU, BUF_U_IN6, &dU);
    x.in((const long long *)uOff, (size_t)nJobs + 1, BUF_U_IN7, &dUOff);
    if (nVirtForced)
        x.in(nVirtForced, (size_t)nBr, BUF_U_IN8, &dForced);
    x.out(nChg, (size_t)nBr, BUF_U_OUT0, &dNChg);
    x.out(outChgTime, (size_t)nBr * outCap, BUF_U_OUT1, &dOCT);
    x.out(outChgType, (size_t)nBr * outCap, BUF_U_OUT2, &dOCTy);
    x.out(logProb, (size_t)nBr, BUF_U_OUT3, &dLP);
    x.out(status, (size_t)nBr, BUF_U_OUT4, &dSt);
    x.out(nVirt, (size_t)nBr, BUF_U_OUT7, &dNVirt);
    x.out((long long *)uUsed, (size_t)nJobs, BUF_U_OUT5, &dUsed);
    int *dOModel = nullptr, *dOSt = nullptr;
    double *dOT0 = nullptr, *dOT1 = nullptr, *dOCT2 = nullptr, *dOLP = nullptr;
    uint8_t *dOA = nullptr, *dOB = nullptr, *dOCTy2 = nullptr;
    long long *dOOff = nullptr;
    if (old) {
        const size_t n = (size_t)old->n, nc = (size_t)old->chgOff[old->n];
        x.in(old->model, n, BUF_U_LIN0, &dOModel);
        x.in(old->tBottom, n, BUF_U_LIN1, &dOT0);
        x.in(old->tTop, n, BUF_U_LIN2, &dOT1);
        x.in(old->typeBottom, n, BUF_U_LIN3, &dOA);
        x.in(old->typeTop, n, BUF_U_LIN4, &dOB);
        x.in((const long long *)old->chgOff, n + 1, BUF_U_LIN5, &dOOff);
        x.in(old->chgTime, nc, BUF_U_LIN6, &dOCT2);
        x.in(old->chgType, nc, BUF_U_LIN7, &dOCTy2);
        x.out(old->logProb, n, BUF_U_LOUT0, &dOLP);
        x.out(old->status, n, BUF_U_LOUT1, &dOSt);
    }
    {
        int rc = x.prepare();
        if (rc == MTT_OK)
            rc = x.upload();
        if (rc != MTT_OK)
            return rc;
    }
    bool oldInFlight = false;
    if (old && old->n > 0) {
        // the probability of the current typing needs no random numbers: it runs on the second
        // stream, beside the matrix exponentials and the sampling of the new typing
        if (!ctx->copyStream)
            MTT_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
        while (ctx->events.size() < 2) {
            cudaEvent_t ev;
            MTT_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            ctx->events.push_back(ev);
        }
        MTT_CUDA(ctx, cudaEventRecord(ctx->events[0], s));
        MTT_CUDA(ctx, cudaStreamWaitEvent(ctx->copyStream, ctx->events[0], 0));
        MTT_CUDA(ctx, mtt_fn_smem(ctx, k3_branch_logp, expm_smem(d)));
        k3_branch_logp<<<(unsigned)((old->n + U_WARPS - 1) / U_WARPS), U_WARPS * 32,
                         expm_smem(d), ctx->copyStream>>>(
            old->n, d, (const double *)ctx->bufs[BUF_U_Q].p,
            (const double *)ctx->bufs[BUF_U_LOGQ].p, dOModel, dOT0, dOT1, dOA, dOB, dOOff,
            dOCT2, dOCTy2, dOLP, dOSt);
        MTT_CHECK_LAUNCH(ctx);
        MTT_CUDA(ctx, cudaEventRecord(ctx->events[1], ctx->copyStream));
        oldInFlight = true;
    }

    // the branches' matrix exponentials, all at once
    double *dPba = (double *)mtt_buf(ctx, BUF_U_PBA, (size_t)nBr * 8);
    if (!dPba)
        return MTT_ERR_CUDA;
    MTT_CUDA(ctx, mtt_fn_smem(ctx, k3_pba, expm_smem(d)));
    k3_pba<<<(unsigned)((nBr + U_WARPS - 1) / U_WARPS), U_WARPS * 32, expm_smem(d), s>>>(
        nBr, d, (const double *)ctx->bufs[BUF_U_Q].p, dModel, dT0, dT1, dA, dB, dPba);
    MTT_CHECK_LAUNCH(ctx);

    const unsigned blocks = (unsigned)((nJobs + U_WARPS - 1) / U_WARPS);
    std::vector<int32_t> hSt(x.staged() ? 0 : nBr);
    for (int attempt = 0; attempt < 12; attempt++) {
        int rc = ensure_powers(ctx, K);
        if (rc != MTT_OK)
            return rc;
        // shared-memory stage sized for the expected number of virtual events (more resident
        // warps for short branches); longer draws use the global scratch below
        int tCap = 32;
        while (tCap < K && tCap < U_TCAP)
            tCap <<= 1;
        const size_t smem = (size_t)U_WARPS * (8 * dd + tCap + tCap / 8) * 8;
        MTT_CUDA(ctx, mtt_fn_smem(ctx, k3_retype, smem));
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
        p.logQ = (const double *)ctx->bufs[BUF_U_LOGQ].p;
        p.mu = (const double *)ctx->bufs[BUF_U_MU].p; p.pow = (const double *)ctx->bufs[BUF_U_POW].p;
        p.logPow = (const double *)ctx->bufs[BUF_U_LOGPOW].p;
        p.logN = (const double *)ctx->bufs[BUF_U_LOGN].p;
        p.meta = (const int *)ctx->bufs[BUF_U_META].p;
        p.brOff = dBrOff; p.model = dModel; p.tBottom = dT0; p.tTop = dT1; p.typeBottom = dA;
        p.typeTop = dB; p.nVirtForced = dForced; p.uniforms = dU; p.uOff = dUOff;
        p.nChg = dNChg; p.outChgTime = dOCT; p.outChgType = dOCTy; p.logProb = dLP;
        p.status = dSt; p.nVirt = dNVirt; p.uUsed = dUsed;
        p.scratchT = scrT; p.scratchTy = scrTy; p.scratchCap = scratchCap; p.tCap = tCap;
        p.pba = dPba;
        k3_retype<<<blocks, U_WARPS * 32, smem, s>>>(p);
        MTT_CHECK_LAUNCH(ctx);
        if (oldInFlight) {              // its results travel with the first download
            MTT_CUDA(ctx, cudaStreamWaitEvent(s, ctx->events[1], 0));
            oldInFlight = false;
        }
        // small calls: all outputs come back in one copy and the status is read from it; large
        // calls fetch the status alone first
        const int32_t *stHost = hSt.data();
        if (x.staged()) {
            rc = x.download();
            if (rc != MTT_OK)
                return rc;
            stHost = status;
        } else {
            MTT_CUDA(ctx, cudaMemcpyAsync(hSt.data(), dSt, nBr * 4, cudaMemcpyDeviceToHost, s));
            MTT_CUDA(ctx, cudaStreamSynchronize(s));
        }
        bool more = false;
        for (int64_t i = 0; i < nBr; i++)
            if (stHost[i] == MTT_ST_NEED_MORE_POWERS)
                more = true;
        if (!more)
            break;
        K = 2 * K + 64;      // the table was too short for some draw: extend and redo the batch
        if (attempt == 11)
            return mtt_fail(ctx, MTT_ERR_TOO_LARGE, "R^k table would exceed %d powers", K);
    }
    if (x.staged())
        return MTT_OK;                 // everything arrived with the last download
    return x.download();
}

}  // namespace

extern "C" {

int mtt_unif_retype_batch(mtt_ctx *ctx, int64_t nJobs, const int64_t *brOff,
                          const int32_t *model, const double *tBottom, const double *tTop,
                          const uint8_t *typeBottom, const uint8_t *typeTop,
                          const int32_t *nVirtForced, const double *uniforms,
                          const int64_t *uOff, int outCap, int32_t *nChg, double *outChgTime,
                          uint8_t *outChgType, double *logProb, int32_t *status, int32_t *nVirt,
                          int64_t *uUsed)
{
    return retype_impl(ctx, nJobs, brOff, model, tBottom, tTop, typeBottom, typeTop, nVirtForced,
                       uniforms, uOff, outCap, nChg, outChgTime, outChgType, logProb, status,
                       nVirt, uUsed, nullptr);
}

int mtt_unif_node_retype_batch(mtt_ctx *ctx, int64_t nOld, const int32_t *modelOld,
                               const double *tBottomOld, const double *tTopOld,
                               const uint8_t *typeBottomOld, const uint8_t *typeTopOld,
                               const int64_t *chgOff, const double *chgTime,
                               const uint8_t *chgType, double *logProbOld, int32_t *statusOld,
                               int64_t nJobs, const int64_t *brOff, const int32_t *model,
                               const double *tBottom, const double *tTop,
                               const uint8_t *typeBottom, const uint8_t *typeTop,
                               const int32_t *nVirtForced, const double *uniforms,
                               const int64_t *uOff, int outCap, int32_t *nChg,
                               double *outChgTime, uint8_t *outChgType, double *logProb,
                               int32_t *status, int32_t *nVirt, int64_t *uUsed)
{
    if (!ctx)
        return MTT_ERR_INVALID_ARG;
    if (nOld < 0 || (nOld > 0 && (!modelOld || !tBottomOld || !tTopOld || !typeBottomOld ||
                                  !typeTopOld || !chgOff || !logProbOld || !statusOld)))
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "mtt_unif_node_retype_batch: NULL argument");
    if (nJobs <= 0 || nOld == 0) {      // nothing to overlap: the two plain calls
        int rc = MTT_OK;
        if (nOld > 0)
            rc = mtt_unif_branch_logp_batch(ctx, nOld, modelOld, tBottomOld, tTopOld,
                                            typeBottomOld, typeTopOld, chgOff, chgTime, chgType,
                                            logProbOld, statusOld);
        if (rc == MTT_OK && nJobs > 0)
            rc = mtt_unif_retype_batch(ctx, nJobs, brOff, model, tBottom, tTop, typeBottom,
                                       typeTop, nVirtForced, uniforms, uOff, outCap, nChg,
                                       outChgTime, outChgType, logProb, status, nVirt, uUsed);
        return rc;
    }
    const OldTyping old{nOld, modelOld, tBottomOld, tTopOld, typeBottomOld, typeTopOld, chgOff,
                        chgTime, chgType, logProbOld, statusOld};
    return retype_impl(ctx, nJobs, brOff, model, tBottom, tTop, typeBottom, typeTop, nVirtForced,
                       uniforms, uOff, outCap, nChg, outChgTime, outChgType, logProb, status,
                       nVirt, uUsed, &old);
}

}  // extern "C"
