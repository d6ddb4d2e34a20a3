// k2_peel.cuh -- pieces shared by the peeling kernels (k2_peel.cu: full re-evaluation of
// candidate trees; k2_update.cu: resident chains with incremental updates)
#pragma once
#include "mtt_common.cuh"

struct mtt_peel_chains;
void k2_chains_release(mtt_peel_chains *s);      // k2_update.cu

struct mtt_peel {
    mtt_ctx *ctx = nullptr;
    int nTips = 0, nCat = 0;
    long long nPatterns = 0, pStride = 0;      // pStride: padded row length of tipStates
    uint8_t *dStates = nullptr;                // [nTips][pStride]
    double *dWeights = nullptr;                // [pStride] (0 beyond nPatterns)
    uint8_t *dConst = nullptr;                 // [pStride] bit k: pattern constant for state k
    double *dSite = nullptr;                   // catRates[C] | catProps[C] | freqs[4]
    double pInv = 0.0, branchRate = 1.0, kappa = 1.0;
    bool haveStates = false, haveWeights = false, haveSite = false, haveHky = false;
    long long version = 0;                     // bumped by every setter (resident chains go stale)
    std::vector<mtt_peel_chains *> chains;     // resident chains built on this handle
    // per-call scratch (grow only)
    void *dOps = nullptr; size_t opsCap = 0;
    double *dMats = nullptr; size_t matsCap = 0;
    double *dPart = nullptr; size_t partCap = 0;
    double *dCold = nullptr; size_t coldCap = 0;   // parked vectors beyond the shared-memory slots
    int *dParent = nullptr; size_t parentCap = 0;
    double *dHeight = nullptr; size_t heightCap = 0;
    double *dMatsIn = nullptr; size_t matsInCap = 0;
    double *dLogL = nullptr; size_t logLCap = 0;
    int *dInfo = nullptr; size_t infoCap = 0;   // per tree: max slots used (or -1 = bad tree)
    double *dTreeParams = nullptr; size_t treeParamsCap = 0;   // per tree (kappa, branchRate)
    double *dNodeRates = nullptr; size_t nodeRatesCap = 0;     // per branch clock rates
    int *dMaxSlots = nullptr;                   // max over the chunk of the above
    int *hMaxSlots = nullptr;                   // pinned host copy
};

// pattern padding granule (largest tile: 128 threads x 2)
constexpr int MTT_PEEL_TP = 256;
constexpr double SCALING_THRESHOLD = 1.0E-100;     // BeerLikelihoodCore.scalingThreshold

// grow-only device scratch of a handle
template <typename T>
inline int grow(mtt_ctx *ctx, T *&ptr, size_t &cap, size_t bytes)
{
    if (bytes + 256 <= cap)
        return MTT_OK;
    if (ptr) {
        MTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
    size_t want = bytes + bytes / 8 + 4096;
    MTT_CUDA(ctx, cudaMalloc((void **)&ptr, want));
    cap = want;
    return MTT_OK;
}

// k2_peel.cu: fixed-order sum of the tile results [B][nTiles][3] and the choice between the
// unscaled and the rescaled evaluation; logL and/or parts[3][partsStride]
int k2_run_finalize(mtt_peel *h, long long B, int nTiles, const double *partial, double *logL,
                    int *mode, double *parts, long long partsStride);
int k2_check_ready(mtt_peel *h, bool needHky);

#ifdef __CUDACC__
// ---- HKY.getTransitionProbabilities [EXT] (closed form) ---------------
__device__ __forceinline__ void hky_matrix(double kappa, const double *f, double distance,
                                           double *matrix)
{
    const double freqA = f[0], freqC = f[1], freqG = f[2], freqT = f[3];
    const double freqR = freqA + freqG, freqY = freqC + freqT;
    const double tab1A = freqA * ((1 / freqR) - 1), tab3A = freqA / freqR, tab2A = 1 - tab3A;
    const double tab1C = freqC * ((1 / freqY) - 1), tab3C = freqC / freqY, tab2C = 1 - tab3C;
    const double tab1G = freqG * ((1 / freqR) - 1), tab3G = tab2A, tab2G = tab3A;
    const double tab1T = freqT * ((1 / freqY) - 1), tab3T = tab2C, tab2T = tab3C;
    const double beta = -1.0 / (2.0 * (freqR * freqY + kappa * (freqA * freqG + freqC * freqT)));
    const double A_R = 1.0 + freqR * (kappa - 1), A_Y = 1.0 + freqY * (kappa - 1);
    const double xx = beta * distance;
    const double bbR = exp(xx * A_R), bbY = exp(xx * A_Y), aa = exp(xx);
    const double oneminusa = 1 - aa;
    const double t1Aaa = tab1A * aa, t1Gaa = tab1G * aa, t1Caa = tab1C * aa, t1Taa = tab1T * aa;
    matrix[0] = freqA + t1Aaa + (tab2A * bbR);
    matrix[1] = freqC * oneminusa;
    matrix[2] = freqG + t1Gaa - (tab3G * bbR);
    matrix[3] = freqT * oneminusa;
    matrix[4] = freqA * oneminusa;
    matrix[5] = freqC + t1Caa + (tab2C * bbY);
    matrix[6] = freqG * oneminusa;
    matrix[7] = freqT + t1Taa - (tab3T * bbY);
    matrix[8] = freqA + t1Aaa - (tab3A * bbR);
    matrix[9] = matrix[1];
    matrix[10] = freqG + t1Gaa + (tab2G * bbR);
    matrix[11] = matrix[3];
    matrix[12] = matrix[4];
    matrix[13] = freqC + t1Caa - (tab3C * bbY);
    matrix[14] = matrix[6];
    matrix[15] = freqT + t1Taa + (tab2T * bbY);
}

// one child's contribution for PPT patterns of one category:
//   a[u][i] = sum_j M[i][j] x[u][j]      (internal child; x from registers or its slot)
//   a[u][i] = M[i][state_u] or 1         (tip child)
// MS = distance in doubles between consecutive double2 of the matrix (2, or 2 C when the
// categories' matrices are interleaved).
template <int PPT, int MS>
__device__ __forceinline__ void child_term(bool isTip, const double *M, const double (&x)[PPT][4],
                                           const int (&tip)[PPT], double (&a)[PPT][4])
{
    if (isTip) {
        // M is transposed for tip children: column of the observed state = M[4 s .. 4 s + 3].
        // Gaps / ambiguity codes (state >= 4) contribute 1; a warp without any takes the
        // select-free path (isTip is uniform over the CTA, so the vote is convergent).
        bool unknown = false;
#pragma unroll
        for (int u = 0; u < PPT; u++)
            unknown |= tip[u] >= 4;
        if (!__any_sync(FULL_MASK, unknown)) {
#pragma unroll
            for (int u = 0; u < PPT; u++) {
                const double *col = M + tip[u] * 2 * MS;
                const double2 c01 = *reinterpret_cast<const double2 *>(col);
                const double2 c23 = *reinterpret_cast<const double2 *>(col + MS);
                a[u][0] = c01.x;
                a[u][1] = c01.y;
                a[u][2] = c23.x;
                a[u][3] = c23.y;
            }
        } else {
#pragma unroll
            for (int u = 0; u < PPT; u++) {
                const bool known = tip[u] < 4;                      // gap / ambiguity: unknown
                const double *col = M + (known ? tip[u] * 2 * MS : 0);
                const double2 c01 = *reinterpret_cast<const double2 *>(col);
                const double2 c23 = *reinterpret_cast<const double2 *>(col + MS);
                a[u][0] = known ? c01.x : 1.0;
                a[u][1] = known ? c01.y : 1.0;
                a[u][2] = known ? c23.x : 1.0;
                a[u][3] = known ? c23.y : 1.0;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const double2 m01 = *reinterpret_cast<const double2 *>(M + i * 2 * MS);
            const double2 m23 = *reinterpret_cast<const double2 *>(M + i * 2 * MS + MS);
#pragma unroll
            for (int u = 0; u < PPT; u++) {
                double s = m01.x * x[u][0];
                s += m01.y * x[u][1];
                s += m23.x * x[u][2];
                s += m23.y * x[u][3];
                a[u][i] = s;
            }
        }
    }
}

// root of one pattern: frequencies, proportion of invariant sites, log.
// Proportion invariant (TreeLikelihood.calculateLogP [EXT]): pInv is added to the root
// partials of constant patterns.  BEAST evaluates WITHOUT rescaling first and only
// repeats the evaluation with rescaling when that returned -infinity; in the rescaled
// pass it still adds pInv to the (now rescaled) root partials.  Both values are formed:
//   llU  the unscaled pass, computed safely: pInv enters at the root's scale, e^-logScale
//   llQ  the rescaled pass, literally (root + pInv, then + logScale)
// together with "would the unscaled pass have underflowed to log(0) here"; k2_finalize
// picks the sum BEAST would have returned.  The two agree whenever no rescaling
// happened on the way to the root (logScale == 0) or pInv == 0.
__device__ __forceinline__ void root_pattern_logl(const double (&root)[4], double logScale,
                                                  double pInv, const uint8_t *constMask,
                                                  long long p, const double *freqs, double &llU,
                                                  double &llQ, bool &under)
{
    double sumU = 0.0, sumQ = 0.0, offsU = logScale;
    bool hasInv = false;
    if (pInv > 0.0) {
        const int cm = constMask[p];
        hasInv = cm != 0;
        // pInv at the root's scale; far below the double range the root is scaled down
        // instead (it underflows gracefully, as it does in BEAST's unscaled pass)
        double invU = pInv, rootMul = 1.0;
        if (hasInv && offsU != 0.0) {
            if (offsU > -300.0) {
                invU = pInv * exp(-offsU);
            } else {
                rootMul = exp(offsU);
                offsU = 0.0;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const bool on = cm & (1 << i);
            const double r = root[i] * rootMul;
            sumU += freqs[i] * (on ? r + invU : r);
            sumQ += freqs[i] * (on ? root[i] + pInv : root[i]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++)
            sumU += freqs[i] * root[i];
        sumQ = sumU;
    }
    llU = log(sumU) + offsU;
    llQ = (sumQ == sumU && offsU == logScale) ? llU : log(sumQ) + logScale;
    // log of the smallest subnormal: below it the unscaled root sum is exactly 0
    under = !hasInv && !(llU >= -744.4400719213812);
}

#endif  // __CUDACC__
