// k2_peel.cu -- piece 2: batched Felsenstein peeling over site patterns on sm_100a.
//
// Replaces BEAST 2 core's TreeLikelihood.calculateLogP + BeerLikelihoodCore (wired by the
// reference at examples/structuredCoalescent.xml:568-572, fxtemplates/MultiTypeTree.xml:184-197)
// for 4-state data: per-branch transition matrices, the three pruning cases (states x states,
// states x partials, partials x partials), per-node rescaling, root integration over rate
// categories and frequencies, and the weighted sum of pattern log-likelihoods.
//
// Design (full re-evaluation of B candidate trees that share one alignment):
//   k2_prepare  one CTA per tree.  Builds child lists from parent[], orders the post-order
//               traversal Sethi-Ullman style (the child needing more live partials first) so
//               that at most ~log2(n)+1 partial vectors are ever live, numbers the slots of
//               parked vectors from the leaves up (small subtrees use slots 0, 1), and writes
//               the traversal as a list of ops together with the two child transition matrices
//               of every op IN OP ORDER (HKY closed form evaluated here from the node heights,
//               or copied from caller matrices).
//   k2_peel     grid = (pattern tiles, trees).  A CTA owns a tile of patterns of one tree and
//               runs the WHOLE traversal for it: the previous op's result stays in registers,
//               parked partial vectors are thread-private columns (no synchronisation between
//               ops) -- slots 0-1 in shared memory, the rarely used deeper ones in an
//               L2-resident global scratch -- op lists and transition matrices stream through
//               shared memory in stages with 1-D bulk async copies (TMA engine, mbarrier
//               completion), tips are read as 1-byte states.  Partials make no round trip
//               through HBM: traffic is tip states + the op/matrix stream, and the kernel is
//               bound by instruction issue / the FP64 pipe instead of HBM.  Rate categories
//               (2, 4, 8) sit in adjacent lanes, one category per thread.
//   k2_finalize fixed-order sum of the per-tile partial sums -> logL[tree].
// fp64 throughout; rescaling follows BeerLikelihoodCore.scalePartials (threshold 1e-100,
// always on -- results agree with the unscaled path to rounding whenever that path is finite).
#include "k2_peel.cuh"

#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

namespace {

constexpr int TP = MTT_PEEL_TP;
constexpr int PREP_THREADS = 256;
constexpr int MAX_SLOTS = 24;

struct PeelOp {
    int c1, c2;     // child refs: >= 0 slot of an internal child, < 0 tip index as -(tip+1)
    int dst;        // slot receiving this node's partials
    int flags;      // bit 0 / 1: child 1 / 2 is the result of the previous op (still in
                    // registers); bit 2: this result must be parked in its slot
};
constexpr int OP_C1_PREV = 1, OP_STORE = 4;

// ---- k2_prepare ------------------------------------------------------------------------
// One CTA per tree; every step is parallel over nodes (level-synchronous sweeps; the number of
// rounds is the depth of the tree):
//   bottom-up  isize(v) = internal nodes in the subtree, need(v) = Sethi-Ullman label + which
//              child is evaluated first (the one needing more live partial vectors)
//   top-down   istart(v) = index of the first op of the subtree
//   emission   op index of v = istart(v) + isize(v) - 1 (post-order: first subtree, second
//              subtree, node); a first child whose sibling is internal parks its result in
//              slot need(sibling) - 1 -- above everything the sibling's subtree will park, so
//              slots are numbered from the leaves up; matrices of both children written in op
//              order.
// shared: int left[N], right[N]; u16 isize[N], istart[N]; u8 need[N], (spare)[N], state[N]
__global__ void __launch_bounds__(PREP_THREADS)
k2_prepare(int n, int C, const int *parentAll, const double *heightAll, const double *matsInAll,
           const double *site, double branchRate, double kappa, PeelOp *opsAll, double *matsAll,
           int *info, int *maxSlotsAll, int xl, const double *treeParams,
           const double *nodeRatesAll)
{
    extern __shared__ int smi[];
    const int N = 2 * n - 1, nOps = n - 1;
    int *left = smi, *right = smi + N;
    uint16_t *isize = reinterpret_cast<uint16_t *>(smi + 2 * N), *istart = isize + N;
    uint8_t *need = reinterpret_cast<uint8_t *>(istart + N), *state = need + 2 * N;
    __shared__ int sRoot, sBad, sMaxSlots, sProgress;

    const int b = blockIdx.x, tid = threadIdx.x;
    const int *parent = parentAll + (size_t)b * N;
    PeelOp *ops = opsAll + (size_t)b * nOps;
    double *mats = matsAll + (size_t)b * nOps * 2 * C * 16;

    if (tid == 0) {
        sRoot = -1;
        sBad = 0;
        sMaxSlots = 0;
    }
    for (int i = tid; i < N; i += PREP_THREADS) {
        left[i] = 0x7fffffff;
        right[i] = -1;
        isize[i] = 0;
        need[i] = 0;
        state[i] = (i < n) ? 1 : 0;        // bottom-up: tips are done
    }
    __syncthreads();
    for (int i = tid; i < N; i += PREP_THREADS) {
        const int p = parent[i];
        if (p < 0) {
            atomicMax(&sRoot, i);
        } else if (p >= N || p < n) {
            sBad = 1;                      // parent out of range or a tip used as a parent
        } else {
            atomicMin(&left[p], i);        // left = lower Nr, right = higher Nr
            atomicMax(&right[p], i);
        }
    }
    __syncthreads();
    for (int i = n + tid; i < N; i += PREP_THREADS)
        if (left[i] == 0x7fffffff || right[i] < 0 || left[i] == right[i])
            sBad = 1;                      // an internal node without exactly two children
    __syncthreads();
    if (sBad || sRoot < n) {
        if (tid == 0)
            info[b] = -1;
        return;
    }
    const int root = sRoot;

    // ---- bottom-up sweep
    for (int round = 0; round < N; round++) {
        __syncthreads();
        if (tid == 0)
            sProgress = 0;
        __syncthreads();
        if (state[root])
            break;
        bool progressed = false;
        for (int v = n + tid; v < N; v += PREP_THREADS) {
            if (state[v])
                continue;
            const int l = left[v], r = right[v];
            if (state[l] != 1 || state[r] != 1)
                continue;
            const int a = need[l] & 0x7f, c = need[r] & 0x7f;
            const int holdL = l >= n, holdR = r >= n;
            const int lf = max(max(a, holdL + c), 1);   // left first
            const int rf = max(max(c, holdR + a), 1);   // right first
            need[v] = (uint8_t)((rf < lf) ? (rf | 0x80) : lf);
            isize[v] = (uint16_t)(isize[l] + isize[r] + 1);
            state[v] = 2;                  // done, visible to parents from the next round on
            progressed = true;
        }
        if (progressed)
            sProgress = 1;
        __syncthreads();
        for (int v = n + tid; v < N; v += PREP_THREADS)
            if (state[v] == 2)
                state[v] = 1;
        if (!sProgress)
            break;                         // a cycle / disconnected piece: never completes
    }
    __syncthreads();
    if (!state[root] || isize[root] != nOps) {
        if (tid == 0)
            info[b] = -1;                  // not a single connected binary tree
        return;
    }

    // ---- top-down sweep (state: 1 = not yet placed, 3 = placed)
    if (tid == 0) {
        istart[root] = 0;
        state[root] = 3;
    }
    for (int round = 0; round < N; round++) {
        __syncthreads();
        if (tid == 0)
            sProgress = 0;
        __syncthreads();
        bool progressed = false;
        for (int v = tid; v < N; v += PREP_THREADS) {
            if (state[v] == 3 || v == root)
                continue;
            const int pp = parent[v];
            if (state[pp] != 3)
                continue;
            const bool pRightFirst = need[pp] & 0x80;
            const int first = pRightFirst ? right[pp] : left[pp];
            istart[v] = (v == first) ? istart[pp] : (uint16_t)(istart[pp] + isize[first]);
            state[v] = 4;
            progressed = true;
        }
        if (progressed)
            sProgress = 1;
        __syncthreads();
        for (int v = tid; v < N; v += PREP_THREADS)
            if (state[v] == 4)
                state[v] = 3;
        if (!sProgress)
            break;
    }
    __syncthreads();

    // ---- emission: ops and, in op order, the transition matrices of both children
    // mats[op][child j][category c][16]
    const double *catRates = site, *freqs = site + 2 * C;
    if (treeParams) {                       // this tree's own (kappa, clock rate)
        kappa = treeParams[2 * (size_t)b];
        branchRate = treeParams[2 * (size_t)b + 1];
    }
    const double *height = heightAll ? heightAll + (size_t)b * N : nullptr;
    const double *matsIn = matsInAll ? matsInAll + (size_t)b * N * C * 16 : nullptr;
    int myMax = 0;
    for (int v = n + tid; v < N; v += PREP_THREADS) {
        const bool rightFirst = need[v] & 0x80;
        const int c1 = rightFirst ? right[v] : left[v];
        const int c2 = rightFirst ? left[v] : right[v];
        PeelOp op;
        // post-order: an internal second child is always the previous op; an internal first
        // child is the previous op when the second child is a tip; otherwise it was parked.
        // The op lists the previous-op child (if any) as its child 1, so that the kernel reads
        // child 1 from registers and child 2 from a tip row or a parked slot -- never both ways.
        const bool swapKids = c2 >= n;
        const int ka = swapKids ? c2 : c1, kb = swapKids ? c1 : c2;
        op.flags = (ka >= n) ? OP_C1_PREV : 0;
        op.c1 = (ka < n) ? -(ka + 1) : 0;
        // a parked first child sits in slot (live vectors the second child needs) - 1: slots
        // are numbered from the leaves up, so the small subtrees -- most of them -- use slots 0, 1
        op.c2 = (kb < n) ? -(kb + 1) : (int)(need[c2] & 0x7f) - 1;
        op.dst = 0;
        if (v != root) {
            const int pp = parent[v];
            const bool pRightFirst = need[pp] & 0x80;
            const int pFirst = pRightFirst ? right[pp] : left[pp];
            const int pSecond = pRightFirst ? left[pp] : right[pp];
            if (pFirst == v && pSecond >= n) {      // my result waits for the sibling subtree
                op.dst = (int)(need[pSecond] & 0x7f) - 1;
                op.flags |= OP_STORE;
                myMax = max(myMax, (int)op.dst + 1);
            }
        }
        ops[istart[v] + isize[v] - 1] = op;
    }
    if (myMax)
        atomicMax(&sMaxSlots, myMax);
    for (int idx = tid; idx < nOps * 2 * C; idx += PREP_THREADS) {
        const int v = n + idx / (2 * C), j = (idx / C) & 1, c = idx % C;
        const int k = istart[v] + isize[v] - 1;
        const bool rightFirst = need[v] & 0x80;
        const int first = rightFirst ? right[v] : left[v], second = rightFirst ? left[v] : right[v];
        const bool swapKids = second >= n;                  // same rule as for the op above
        const int child = ((j == 0) != swapKids) ? first : second;
        double *M = mats + ((size_t)(k * 2 + j) * C + c) * 16;
        double T[16];
        if (matsIn) {
            const double *src = matsIn + ((size_t)child * C + c) * 16;
            for (int e = 0; e < 16; e++)
                T[e] = src[e];
        } else {
            // TreeLikelihood.traverse [EXT]: jointBranchRate = categoryRate * branchRate, the
            // branch rate being branchRateModel.getRateForBranch(node): one number under a
            // strict clock, one per branch under a relaxed one
            const double br = nodeRatesAll ? nodeRatesAll[(size_t)b * N + child] : branchRate;
            const double jointBranchRate = catRates[c] * br;
            const double distance = (height[v] - height[child]) * jointBranchRate;
            hky_matrix(kappa, freqs, distance, T);
        }
        // internal child: row-major P[i][j] (row = parent state) for the mat-vec; tip child:
        // TRANSPOSED, so that the column of the observed state is 4 contiguous doubles
        // xl (categories across lanes): the 8 double2 of the C matrices of one child are
        // interleaved, [pair][category], so that the lanes of a pattern read adjacent 16-byte
        // words
        const bool tip = child < n;
        for (int e = 0; e < 16; e++) {
            const double v = tip ? T[(e & 3) * 4 + (e >> 2)] : T[e];
            if (xl)
                mats[(size_t)(k * 2 + j) * C * 16 + ((size_t)(e >> 1) * C + c) * 2 + (e & 1)] = v;
            else
                M[e] = v;
        }
    }
    __syncthreads();
    if (tid == 0) {
        info[b] = sMaxSlots;
        atomicMax(maxSlotsAll, sMaxSlots);
    }
}

// ---- k2_peel ---------------------------------------------------------------------------

template <int C>
struct PeelCfg {
    // small stages: shared memory decides how many CTAs (warps) an SM holds, and the kernel
    // is latency-bound on dependent fp64 chains, so residency matters more than stage depth
    static constexpr int OPS_PER_STAGE = (C <= 4) ? 8 : 4;
    static constexpr int MAT_DOUBLES_PER_OP = 2 * C * 16;
    static constexpr int STAGE_BYTES = OPS_PER_STAGE * (16 + MAT_DOUBLES_PER_OP * 8);
};

// rate categories across lanes: C = 2, 4, 8
__host__ __device__ constexpr bool cats_across_lanes(int C)
{
    return C == 2 || C == 4 || C == 8;
}

// TPT threads per CTA, PPT patterns per thread.
// XL = false: a thread owns all C categories of its patterns; a tile is TPT * PPT patterns
//   (pattern of (thread, u) = tile * TPT * PPT + u * TPT + tid, so global reads stay coalesced).
// XL = true (C = 2, 4, 8): the C categories of a pattern sit in C adjacent lanes, each thread
//   carries ONE category -- the register and shared-memory footprint of the C = 1 kernel, so
//   residency does not collapse with the category count; a tile is (TPT / C) * PPT patterns.
//   The per-pattern rescaling maximum and the root sum over categories are lane-group shuffles.
template <int C, int TPT, int PPT, bool XL>
__global__ void __launch_bounds__(TPT)
k2_peel(int n, long long P, long long pStride, const uint8_t *states, const double *weights,
        const uint8_t *constMask, const double *site, double pInv, const PeelOp *opsAll,
        const double *matsAll, const int *info, int nSlots, double *partial, int nTiles,
        double *patternLogL, int smemSlots, double *cold, int treeBase)
{
    using Cfg = PeelCfg<C>;
    extern __shared__ __align__(16) unsigned char smem[];
    // layout: [2 stages: ops | mats][barriers][smemSlots parked vectors of CR*4*PPT*TPT doubles]
    unsigned char *stageBase = smem;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 2 * Cfg::STAGE_BYTES);
    double *stackBase = reinterpret_cast<double *>(smem + 2 * Cfg::STAGE_BYTES + 16);
    __shared__ double red[3 * ((TPT + 31) / 32)];

    constexpr int CR = XL ? 1 : C;            // categories held per thread
    constexpr int LPP = XL ? C : 1;           // lanes per pattern
    constexpr int TPP = TPT / LPP;            // patterns per u-slice of the tile
    constexpr int MS = XL ? 2 * C : 2;        // double2 stride inside a staged matrix
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cat = XL ? (tid & (C - 1)) : 0;
    const int tile = blockIdx.x, b = blockIdx.y + treeBase;
    const int nOps = n - 1;
    const long long p0 = (long long)tile * TPP * PPT + tid / LPP;   // + u * TPP; < pStride
    const PeelOp *ops = opsAll + (size_t)b * nOps;
    const double *mats = matsAll + (size_t)b * nOps * Cfg::MAT_DOUBLES_PER_OP;
    const int treeInfo = info[b];

    if (treeInfo < 0 || treeInfo > nSlots) {                 // malformed tree: NaN, loudly
        if (tid == 0) {
            double *out = partial + ((size_t)b * nTiles + tile) * 3;
            out[0] = out[1] = NAN;
            out[2] = 0.0;
        }
        if (patternLogL)
            for (int u = 0; u < PPT; u++)
                if (p0 + u * TPP < P)
                    patternLogL[p0 + u * TPP] = patternLogL[pStride + p0 + u * TPP] = NAN;
        return;
    }
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int nStages = (nOps + Cfg::OPS_PER_STAGE - 1) / Cfg::OPS_PER_STAGE;
    auto issue = [&](int s) {
        const int k0 = s * Cfg::OPS_PER_STAGE;
        const int cnt = min(Cfg::OPS_PER_STAGE, nOps - k0);
        unsigned char *dst = stageBase + (s & 1) * Cfg::STAGE_BYTES;
        const uint32_t opBytes = cnt * 16, matBytes = cnt * Cfg::MAT_DOUBLES_PER_OP * 8;
        mbar_arrive_expect_tx(&bars[s & 1], opBytes + matBytes);
        bulk_g2s(dst, ops + k0, opBytes, &bars[s & 1]);
        bulk_g2s(dst + Cfg::OPS_PER_STAGE * 16, mats + (size_t)k0 * Cfg::MAT_DOUBLES_PER_OP,
                 matBytes, &bars[s & 1]);
    };
    if (tid == 0)
        issue(0);

    // tip states of the op about to run are fetched one op ahead
    auto load_tips = [&](const PeelOp &op, int (&t1)[PPT], int (&t2)[PPT]) {
        if (op.c1 < 0) {
            const uint8_t *row = states + (size_t)(-op.c1 - 1) * pStride + p0;
#pragma unroll
            for (int u = 0; u < PPT; u++)
                t1[u] = __ldg(row + u * TPP);
        }
        if (op.c2 < 0) {
            const uint8_t *row = states + (size_t)(-op.c2 - 1) * pStride + p0;
#pragma unroll
            for (int u = 0; u < PPT; u++)
                t2[u] = __ldg(row + u * TPP);
        }
    };

    double prev[CR][PPT][4];         // result of the previous op (top of the stack, in registers)
    double logScale[PPT];
#pragma unroll
    for (int u = 0; u < PPT; u++) {
        logScale[u] = 0.0;
#pragma unroll
        for (int c = 0; c < CR; c++)
#pragma unroll
            for (int i = 0; i < 4; i++)
                prev[c][u][i] = 1.0;
    }
    // parked vectors: slots [0, smemSlots) in shared memory, the rarely used deeper ones in this
    // CTA's piece of the global scratch (thread-contiguous, so the accesses coalesce)
    constexpr size_t SLOT_DOUBLES = (size_t)CR * 4 * PPT * TPT;
    double *coldBase = cold + ((size_t)blockIdx.y * gridDim.x + tile) *
                                  (size_t)max(nSlots - smemSlots, 0) * SLOT_DOUBLES;
    auto slot_ptr = [&](int slot) -> double * {
        return (slot < smemSlots) ? stackBase + (size_t)slot * SLOT_DOUBLES
                                  : coldBase + (size_t)(slot - smemSlots) * SLOT_DOUBLES;
    };
    uint32_t phaseBits = 0;
    for (int s = 0; s < nStages; s++) {
        if (tid == 0 && s + 1 < nStages)
            issue(s + 1);      // the other stage was released by the __syncthreads below
        mbar_wait(&bars[s & 1], (phaseBits >> (s & 1)) & 1u);
        phaseBits ^= 1u << (s & 1);
        const unsigned char *st = stageBase + (s & 1) * Cfg::STAGE_BYTES;
        const PeelOp *sOps = reinterpret_cast<const PeelOp *>(st);
        const double *sMats = reinterpret_cast<const double *>(st + Cfg::OPS_PER_STAGE * 16);
        const int cnt = min(Cfg::OPS_PER_STAGE, nOps - s * Cfg::OPS_PER_STAGE);
        int nt1[PPT], nt2[PPT];
#pragma unroll
        for (int u = 0; u < PPT; u++)
            nt1[u] = nt2[u] = 4;
        load_tips(sOps[0], nt1, nt2);
#pragma unroll 1
        for (int k = 0; k < cnt; k++) {
            const PeelOp op = sOps[k];
            int t1[PPT], t2[PPT];
#pragma unroll
            for (int u = 0; u < PPT; u++) {
                t1[u] = nt1[u];
                t2[u] = nt2[u];
            }
            if (k + 1 < cnt)
                load_tips(sOps[k + 1], nt1, nt2);
            int mxhi[PPT];       // partials are >= 0: their order is the order of the high words
#pragma unroll
            for (int u = 0; u < PPT; u++)
                mxhi[u] = 0;
#pragma unroll
            for (int c = 0; c < CR; c++) {
                const double *M1 = sMats + (size_t)k * Cfg::MAT_DOUBLES_PER_OP +
                                   (XL ? cat * 2 : c * 16);
                const double *M2 = M1 + C * 16;
                double x[PPT][4], a[PPT][4], bb[PPT][4];
                // child 1: a tip, or the previous op's result (registers)
                child_term<PPT, MS>(op.c1 < 0, M1, prev[c], t1, a);
                // child 2: a tip, or a parked vector
                if (op.c2 >= 0) {
                    if (op.c2 < smemSlots) {
                        const double *src = stackBase + (size_t)op.c2 * SLOT_DOUBLES +
                                            (size_t)c * 4 * PPT * TPT + tid;
#pragma unroll
                        for (int i = 0; i < 4; i++)
#pragma unroll
                            for (int u = 0; u < PPT; u++)
                                x[u][i] = src[(i * PPT + u) * TPT];
                    } else {
                        const double *src = slot_ptr(op.c2) + (size_t)c * 4 * PPT * TPT + tid;
#pragma unroll
                        for (int i = 0; i < 4; i++)
#pragma unroll
                            for (int u = 0; u < PPT; u++)
                                x[u][i] = __ldcg(src + (i * PPT + u) * TPT);
                    }
                }
                child_term<PPT, MS>(op.c2 < 0, M2, x, t2, bb);
#pragma unroll
                for (int u = 0; u < PPT; u++)
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        prev[c][u][i] = a[u][i] * bb[u][i];
                        mxhi[u] = max(mxhi[u], __double2hiint(prev[c][u][i]));
                    }
            }
            // BeerLikelihoodCore.scalePartials [EXT]: rescale when the largest partial of the
            // pattern (over categories and states) drops below the threshold.  The cheap test
            // on the high word is conservative; the exact maximum is formed only when needed.
#pragma unroll
            for (int u = 0; u < PPT; u++) {
                if (XL) {
                    // maximum over the pattern's categories = over its lane group
#pragma unroll
                    for (int o = 1; o < LPP; o <<= 1)
                        mxhi[u] = max(mxhi[u], __shfl_xor_sync(FULL_MASK, mxhi[u], o));
                }
                const bool low = mxhi[u] <= __double2hiint(SCALING_THRESHOLD);
                // (the shuffles below need the whole warp: the branch is taken warp-wide)
                if (XL ? __any_sync(FULL_MASK, low) : low) {
                    double mx = 0.0;
#pragma unroll
                    for (int c = 0; c < CR; c++)
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            mx = fmax(mx, prev[c][u][i]);
                    if (XL) {
#pragma unroll
                        for (int o = 1; o < LPP; o <<= 1)
                            mx = fmax(mx, __shfl_xor_sync(FULL_MASK, mx, o));
                    }
                    if (low && mx < SCALING_THRESHOLD) {
#pragma unroll
                        for (int c = 0; c < CR; c++)
#pragma unroll
                            for (int i = 0; i < 4; i++)
                                prev[c][u][i] /= mx;
                        logScale[u] += log(mx);
                    }
                }
            }
            if (op.flags & OP_STORE) {
                double *dst = slot_ptr(op.dst) + tid;
#pragma unroll
                for (int c = 0; c < CR; c++)
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int u = 0; u < PPT; u++)
                            dst[((c * 4 + i) * PPT + u) * TPT] = prev[c][u][i];
            }
        }
        __syncthreads();       // every thread is done with this stage's ops and matrices
    }

    // root (= result of the last op): integratePartials, proportion invariant, frequencies, log
    const double *catProps = site + C, *freqs = site + 2 * C;
    double contribU = 0.0, contribQ = 0.0;
    bool anyUnder = false;
#pragma unroll
    for (int u = 0; u < PPT; u++) {
        const long long p = p0 + u * TPP;
        double root[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            double r;
            if (XL) {
                // sum over the lane group in category order 0, 1, .., C-1
                const double mine = prev[0][u][i] * catProps[cat];
                const int base = lane & ~(LPP - 1);
                r = __shfl_sync(FULL_MASK, mine, base);
#pragma unroll
                for (int c = 1; c < LPP; c++)
                    r += __shfl_sync(FULL_MASK, mine, base + c);
            } else {
                r = prev[0][u][i] * catProps[0];
#pragma unroll
                for (int c = 1; c < CR; c++)
                    r += prev[c][u][i] * catProps[c];
            }
            root[i] = r;
        }
        double llU, llQ;
        bool under;
        root_pattern_logl(root, logScale[u], pInv, constMask, p, freqs, llU, llQ, under);
        if (patternLogL && p < P && cat == 0) {
            patternLogL[p] = llU;
            patternLogL[pStride + p] = llQ;
        }
        if (p < P && cat == 0) {
            contribU += llU * weights[p];
            contribQ += llQ * weights[p];
            anyUnder |= under;
        }
    }
    contribU = warp_sum_down(contribU);
    contribQ = warp_sum_down(contribQ);
    anyUnder = __any_sync(FULL_MASK, anyUnder);
    if (lane == 0) {
        red[warp] = contribU;
        red[(TPT + 31) / 32 + warp] = contribQ;
        red[2 * ((TPT + 31) / 32) + warp] = anyUnder ? 1.0 : 0.0;
    }
    __syncthreads();
    if (tid == 0) {
        double tU = 0.0, tQ = 0.0, f = 0.0;
#pragma unroll
        for (int w = 0; w < (TPT + 31) / 32; w++) {
            tU += red[w];
            tQ += red[(TPT + 31) / 32 + w];
            f += red[2 * ((TPT + 31) / 32) + w];
        }
        double *out = partial + ((size_t)b * nTiles + tile) * 3;
        out[0] = tU;
        out[1] = tQ;
        out[2] = f;
    }
}

// Per tree: fixed-order sums of the tile results.  BEAST's TreeLikelihood [EXT] returns the
// unscaled evaluation unless it is -infinity (some pattern underflowed), in which case the
// rescaled evaluation is returned: picked here from the two sums every tile delivers.
// `mode` (optional, per tree): 1 = the rescaled evaluation was picked.
__global__ void k2_finalize(long long B, int nTiles, const double *partial, double *logL,
                            int *mode, double *parts, long long partsStride)
{
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B)
        return;
    double sU = 0.0, sQ = 0.0, f = 0.0;
    for (int t = 0; t < nTiles; t++) {
        const double *in = partial + ((size_t)b * nTiles + t) * 3;
        sU += in[0];
        sQ += in[1];
        f += in[2];
    }
    if (logL)
        logL[b] = (f > 0.0) ? sQ : sU;
    if (mode)
        mode[b] = f > 0.0;
    if (parts) {           // a pattern shard: the pick is made after the sum over shards
        parts[b] = sU;
        parts[partsStride + b] = sQ;
        parts[2 * partsStride + b] = f;
    }
}

// per-pattern output of a single tree: keep the evaluation k2_finalize picked
__global__ void k2_pick_patterns(long long P, long long pStride, const int *mode, double *pat)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P && mode[0])
        pat[p] = pat[pStride + p];
}

// constant-site masks (TreeLikelihood.calcConstantPatternIndices [EXT]): pattern p is constant
// for state k when every unambiguous tip shows k
__global__ void k2_const_mask(int n, long long P, long long pStride, const uint8_t *states,
                              uint8_t *mask)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= pStride)
        return;
    int m = 0xf;
    if (p < P)
        for (int t = 0; t < n; t++) {
            const int s = states[(size_t)t * pStride + p];
            if (s < 4)
                m &= (1 << s);
        }
    mask[p] = (uint8_t)m;
}

template <int C, int TPT, int PPT>
int launch_peel_t(mtt_peel *h, long long B, int nSlots, int smemSlots, double *patternLogL)
{
    mtt_ctx *ctx = h->ctx;
    using Cfg = PeelCfg<C>;
    constexpr bool XL = cats_across_lanes(C);
    constexpr int CR = XL ? 1 : C;
    constexpr size_t SLOT_BYTES = (size_t)CR * 4 * PPT * TPT * 8;
    const int nTiles = (int)(h->pStride / ((TPT / (XL ? C : 1)) * PPT));
    smemSlots = std::min(smemSlots, nSlots);
    const size_t smem = 2 * Cfg::STAGE_BYTES + 16 + (size_t)smemSlots * SLOT_BYTES;
    auto kern = k2_peel<C, TPT, PPT, XL>;
    MTT_CUDA(ctx, mtt_fn_smem(ctx, kern, smem));
    // global scratch for the deeper slots, at most ~1 GB: larger batches run in sub-batches
    const size_t coldPerTree = (size_t)nTiles * (size_t)(nSlots - smemSlots) * SLOT_BYTES;
    long long sub = B;
    if (coldPerTree > 0)
        sub = std::max(1LL, std::min(B, (long long)((1ull << 30) / coldPerTree)));
    int rc = grow(ctx, h->dCold, h->coldCap, (size_t)sub * coldPerTree + 256);
    if (rc != MTT_OK)
        return rc;
    for (long long b0 = 0; b0 < B; b0 += sub) {
        dim3 grid((unsigned)nTiles, (unsigned)std::min(sub, B - b0));
        kern<<<grid, TPT, smem, ctx->stream>>>(h->nTips, h->nPatterns, h->pStride, h->dStates,
                                               h->dWeights, h->dConst, h->dSite, h->pInv,
                                               (const PeelOp *)h->dOps, h->dMats, h->dInfo,
                                               nSlots, h->dPart, nTiles, patternLogL, smemSlots,
                                               h->dCold, (int)b0);
        MTT_CHECK_LAUNCH(ctx);
    }
    return MTT_OK;
}

// Tile shape: 128 threads x 2 patterns when two such CTAs fit an SM (parked partial vectors
// live in shared memory); otherwise 128 x 1, then 64 x 1, then 32 x 1.
template <int C>
int launch_peel(mtt_peel *h, long long B, int nSlots, int *nTilesOut, double *patternLogL)
{
    constexpr int LPP = cats_across_lanes(C) ? C : 1;      // lanes per pattern
    constexpr int CR = cats_across_lanes(C) ? 1 : C;       // categories per thread
    // Shared memory holds the two shallowest slots -- with slots numbered from the leaves up
    // they take ~90 % of the parked traffic of coalescent-shaped trees -- and the deeper ones
    // go to the L2-resident global scratch, so that six CTAs (24 warps) stay resident per SM
    // whatever the depth of the tree.  With CR > 1 categories per thread a slot is CR times
    // larger: fall back to smaller tiles as before.
    constexpr int SMEM_SLOTS = 2;
    if (CR <= 2) {
        *nTilesOut = (int)(h->pStride / (256 / LPP));
        return launch_peel_t<C, 128, 2>(h, B, nSlots, SMEM_SLOTS, patternLogL);
    }
    if (CR <= 4) {
        *nTilesOut = (int)(h->pStride / (128 / LPP));
        return launch_peel_t<C, 128, 1>(h, B, nSlots, SMEM_SLOTS, patternLogL);
    }
    *nTilesOut = (int)(h->pStride / (64 / LPP));
    return launch_peel_t<C, 64, 1>(h, B, nSlots, SMEM_SLOTS, patternLogL);
}

// run the pipeline on device-resident parent/height (or matsIn) for a chunk of trees
int peel_chunk(mtt_peel *h, long long B, const int *dParent, const double *dHeight,
               const double *dMatsIn, double *dLogL, double *patternLogL,
               const double *dTreeParams = nullptr, double *dParts = nullptr,
               long long partsStride = 0, const double *dNodeRates = nullptr)
{
    mtt_ctx *ctx = h->ctx;
    const int n = h->nTips, C = h->nCat, N = 2 * n - 1, nOps = n - 1;
    const int maxTiles = (int)(h->pStride / 4);     // smallest tile: 32 threads / 8 categories
    int nTiles = 0;
    int rc;
    if ((rc = grow(ctx, h->dOps, h->opsCap, (size_t)B * nOps * sizeof(PeelOp))) != MTT_OK)
        return rc;
    if ((rc = grow(ctx, h->dMats, h->matsCap, (size_t)B * nOps * 2 * C * 16 * 8)) != MTT_OK)
        return rc;
    if ((rc = grow(ctx, h->dPart, h->partCap, (size_t)B * maxTiles * 24)) != MTT_OK)
        return rc;
    if ((rc = grow(ctx, h->dInfo, h->infoCap, (size_t)B * 4)) != MTT_OK)
        return rc;
    const size_t prepSmem = (size_t)N * 15 + 64;
    if (prepSmem > 227 * 1024)
        return mtt_fail(ctx, MTT_ERR_TOO_LARGE, "peeling: %d tips exceed the traversal planner's "
                        "shared memory (limit ~7700 tips)", n);
    MTT_CUDA(ctx, mtt_fn_smem(ctx, k2_prepare, prepSmem));
    if (!h->dMaxSlots) {
        MTT_CUDA(ctx, cudaMalloc((void **)&h->dMaxSlots, 256));
        MTT_CUDA(ctx, cudaMallocHost((void **)&h->hMaxSlots, sizeof(int)));
    }
    MTT_CUDA(ctx, cudaMemsetAsync(h->dMaxSlots, 0, sizeof(int), ctx->stream));
    k2_prepare<<<(unsigned)B, PREP_THREADS, prepSmem, ctx->stream>>>(
        n, C, dParent, dHeight, dMatsIn, h->dSite, h->branchRate, h->kappa, (PeelOp *)h->dOps,
        h->dMats, h->dInfo, h->dMaxSlots, cats_across_lanes(C) ? 1 : 0, dTreeParams, dNodeRates);
    MTT_CHECK_LAUNCH(ctx);
    // the number of parked partial vectors sizes the peel kernel's shared memory (a few for
    // coalescent-shaped trees, up to floor(log2 n) for balanced ones): read it back
    MTT_CUDA(ctx, cudaMemcpyAsync(h->hMaxSlots, h->dMaxSlots, sizeof(int), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    MTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int nSlots = std::max(1, std::min(*h->hMaxSlots, MAX_SLOTS));
    switch (C) {
    case 1: rc = launch_peel<1>(h, B, nSlots, &nTiles, patternLogL); break;
    case 2: rc = launch_peel<2>(h, B, nSlots, &nTiles, patternLogL); break;
    case 3: rc = launch_peel<3>(h, B, nSlots, &nTiles, patternLogL); break;
    case 4: rc = launch_peel<4>(h, B, nSlots, &nTiles, patternLogL); break;
    case 5: rc = launch_peel<5>(h, B, nSlots, &nTiles, patternLogL); break;
    case 6: rc = launch_peel<6>(h, B, nSlots, &nTiles, patternLogL); break;
    case 7: rc = launch_peel<7>(h, B, nSlots, &nTiles, patternLogL); break;
    case 8: rc = launch_peel<8>(h, B, nSlots, &nTiles, patternLogL); break;
    default:
        return mtt_fail(ctx, MTT_ERR_TOO_LARGE, "peeling: nCat=%d outside [1,8]", C);
    }
    if (rc != MTT_OK)
        return rc;
    // dInfo is consumed by now: its first word doubles as the picked-evaluation flag of a
    // single-tree per-pattern request
    k2_finalize<<<(unsigned)((B + 127) / 128), 128, 0, ctx->stream>>>(
        B, nTiles, h->dPart, dLogL, patternLogL ? h->dInfo : nullptr, dParts, partsStride);
    MTT_CHECK_LAUNCH(ctx);
    if (patternLogL) {
        k2_pick_patterns<<<(unsigned)((h->nPatterns + 255) / 256), 256, 0, ctx->stream>>>(
            h->nPatterns, h->pStride, h->dInfo, patternLogL);
        MTT_CHECK_LAUNCH(ctx);
    }
    return MTT_OK;
}

long long chunk_trees(const mtt_peel *h)
{
    // bound the op/matrix stream scratch to ~512 MB per chunk
    const size_t perTree = (size_t)(h->nTips - 1) * (sizeof(PeelOp) + 2 * h->nCat * 16 * 8);
    long long c = (long long)((512ull << 20) / perTree);
    return std::max(1LL, std::min(c, 65535LL));
}

int check_ready(mtt_peel *h, bool needHky)
{
    if (!h)
        return MTT_ERR_INVALID_ARG;
    if (!h->haveStates || !h->haveWeights || !h->haveSite || (needHky && !h->haveHky))
        return mtt_fail(h->ctx, MTT_ERR_STATE,
                        "peeling: set_tip_states, set_pattern_weights, set_site_model%s must be "
                        "called first", needHky ? " and set_hky" : "");
    return MTT_OK;
}

}  // namespace

int k2_check_ready(mtt_peel *h, bool needHky)
{
    return check_ready(h, needHky);
}

int k2_run_finalize(mtt_peel *h, long long B, int nTiles, const double *partial, double *logL,
                    int *mode, double *parts, long long partsStride)
{
    mtt_ctx *ctx = h->ctx;
    k2_finalize<<<(unsigned)((B + 127) / 128), 128, 0, ctx->stream>>>(B, nTiles, partial, logL,
                                                                      mode, parts, partsStride);
    MTT_CHECK_LAUNCH(ctx);
    return MTT_OK;
}

extern "C" {

int mtt_peel_create(mtt_ctx *ctx, int nTips, int64_t nPatterns, int nCat, mtt_peel **out)
{
    if (!ctx || !out)
        return MTT_ERR_INVALID_ARG;
    *out = nullptr;
    if (nTips < 2 || nPatterns < 1)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "mtt_peel_create: need >= 2 tips, >= 1 pattern");
    if (nCat < 1 || nCat > 8)
        return mtt_fail(ctx, MTT_ERR_TOO_LARGE, "mtt_peel_create: nCat=%d outside [1,8]", nCat);
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    mtt_peel *h = new mtt_peel();
    h->ctx = ctx;
    h->nTips = nTips;
    h->nCat = nCat;
    h->nPatterns = nPatterns;
    h->pStride = (nPatterns + TP - 1) / TP * TP;
    cudaError_t e = cudaMalloc((void **)&h->dStates, (size_t)nTips * h->pStride + 256);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&h->dWeights, h->pStride * 8 + 256);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&h->dConst, h->pStride + 256);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&h->dSite, (2 * nCat + 4) * 8 + 256);
    if (e != cudaSuccess) {
        mtt_peel_destroy(h);
        return mtt_fail(ctx, MTT_ERR_CUDA, "mtt_peel_create: cudaMalloc failed: %s",
                        cudaGetErrorString(e));
    }
    *out = h;
    return MTT_OK;
}

int mtt_peel_destroy(mtt_peel *h)
{
    if (!h)
        return MTT_OK;
    cudaSetDevice(h->ctx->device);
    cudaStreamSynchronize(h->ctx->stream);
    for (mtt_peel_chains *c : h->chains)
        k2_chains_release(c);               // chains that outlive their handle lose the device side
    void *ptrs[] = {h->dStates, h->dWeights, h->dConst, h->dSite, h->dOps, h->dMats, h->dPart, h->dCold,
                    h->dParent, h->dHeight, h->dMatsIn, h->dLogL, h->dInfo, h->dMaxSlots,
                    h->dTreeParams, h->dNodeRates};
    if (h->hMaxSlots)
        cudaFreeHost(h->hMaxSlots);
    for (void *p : ptrs)
        if (p)
            cudaFree(p);
    delete h;
    return MTT_OK;
}

int mtt_peel_set_tip_states(mtt_peel *h, const int32_t *states)
{
    if (!h || !states)
        return MTT_ERR_INVALID_ARG;
    mtt_ctx *ctx = h->ctx;
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    // int32 [nTips][nPatterns] -> uint8 [nTips][pStride]; padding patterns are "unknown"
    std::vector<uint8_t> packed((size_t)h->nTips * h->pStride, 4);
    for (int t = 0; t < h->nTips; t++)
        for (long long p = 0; p < h->nPatterns; p++) {
            const int32_t s = states[(size_t)t * h->nPatterns + p];
            packed[(size_t)t * h->pStride + p] = (uint8_t)((s >= 0 && s < 4) ? s : 4);
        }
    MTT_CUDA(ctx, cudaMemcpyAsync(h->dStates, packed.data(), packed.size(),
                                  cudaMemcpyHostToDevice, ctx->stream));
    k2_const_mask<<<(unsigned)((h->pStride + 255) / 256), 256, 0, ctx->stream>>>(
        h->nTips, h->nPatterns, h->pStride, h->dStates, h->dConst);
    MTT_CHECK_LAUNCH(ctx);
    MTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    h->haveStates = true;
    h->version++;
    return MTT_OK;
}

int mtt_peel_set_pattern_weights(mtt_peel *h, const int32_t *weights)
{
    if (!h || !weights)
        return MTT_ERR_INVALID_ARG;
    mtt_ctx *ctx = h->ctx;
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<double> w((size_t)h->pStride, 0.0);
    for (long long p = 0; p < h->nPatterns; p++)
        w[p] = (double)weights[p];
    MTT_CUDA(ctx, cudaMemcpyAsync(h->dWeights, w.data(), w.size() * 8, cudaMemcpyHostToDevice,
                                  ctx->stream));
    MTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    h->haveWeights = true;
    h->version++;
    return MTT_OK;
}

int mtt_peel_set_site_model(mtt_peel *h, const double *catRates, const double *catProps,
                            const double *freqs, double pInv, double branchRate)
{
    if (!h || !catRates || !catProps || !freqs)
        return MTT_ERR_INVALID_ARG;
    mtt_ctx *ctx = h->ctx;
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<double> s(2 * h->nCat + 4);
    for (int c = 0; c < h->nCat; c++) {
        s[c] = catRates[c];
        s[h->nCat + c] = catProps[c];
    }
    for (int i = 0; i < 4; i++)
        s[2 * h->nCat + i] = freqs[i];
    MTT_CUDA(ctx, cudaMemcpyAsync(h->dSite, s.data(), s.size() * 8, cudaMemcpyHostToDevice,
                                  ctx->stream));
    MTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    h->pInv = pInv;
    h->branchRate = branchRate;
    h->haveSite = true;
    h->version++;
    return MTT_OK;
}

int mtt_peel_set_hky(mtt_peel *h, double kappa)
{
    if (!h)
        return MTT_ERR_INVALID_ARG;
    h->kappa = kappa;
    h->haveHky = true;
    h->version++;
    return MTT_OK;
}

int mtt_peel_logl_batch_dev(mtt_peel *h, int64_t B, const int32_t *parent, const double *height,
                            double *logL)
{
    int rc = check_ready(h, true);
    if (rc != MTT_OK)
        return rc;
    mtt_ctx *ctx = h->ctx;
    if (B < 0 || !parent || !height || !logL)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "mtt_peel_logl_batch_dev: bad arguments");
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    const int N = 2 * h->nTips - 1;
    const long long chunk = chunk_trees(h);
    for (long long b0 = 0; b0 < B; b0 += chunk) {
        const long long nb = std::min(chunk, (long long)B - b0);
        rc = peel_chunk(h, nb, parent + b0 * N, height + b0 * N, nullptr, logL + b0, nullptr);
        if (rc != MTT_OK)
            return rc;
    }
    return MTT_OK;
}

int mtt_peel_logl_parts_batch_dev(mtt_peel *h, int64_t B, const int32_t *parent,
                                  const double *height, double *parts)
{
    int rc = check_ready(h, true);
    if (rc != MTT_OK)
        return rc;
    mtt_ctx *ctx = h->ctx;
    if (B < 0 || !parent || !height || !parts)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "mtt_peel_logl_parts_batch_dev: bad arguments");
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    const int N = 2 * h->nTips - 1;
    const long long chunk = chunk_trees(h);
    for (long long b0 = 0; b0 < B; b0 += chunk) {
        const long long nb = std::min(chunk, (long long)B - b0);
        rc = peel_chunk(h, nb, parent + b0 * N, height + b0 * N, nullptr, nullptr, nullptr,
                        nullptr, parts + b0, (long long)B);
        if (rc != MTT_OK)
            return rc;
    }
    return MTT_OK;
}

static int peel_host(mtt_peel *h, int64_t B, const int32_t *parent, const double *height,
                     const double *mats, double *logL, double *patternLogL,
                     const double *kappaPerTree = nullptr, const double *ratePerTree = nullptr,
                     double *partsHost = nullptr, const double *nodeRates = nullptr)
{
    mtt_ctx *ctx = h->ctx;
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int N = 2 * h->nTips - 1, C = h->nCat;
    const long long chunk = chunk_trees(h);
    int rc;
    // dLogL: logL[B], followed by parts[3][B] when the caller wants the two sums separately
    if ((rc = grow(ctx, h->dLogL, h->logLCap, (size_t)std::max<int64_t>(B, 1) * 8 * 4)) != MTT_OK)
        return rc;
    double *dParts = partsHost ? h->dLogL + B : nullptr;
    double *dPat = nullptr;
    if (patternLogL) {
        // reuse the weights-sized scratch behind dPart for the per-pattern output
        if ((rc = grow(ctx, h->dHeight, h->heightCap,
                       std::max((size_t)N * 8, (size_t)h->pStride * 16))) != MTT_OK)
            return rc;
    }
    for (long long b0 = 0; b0 < B; b0 += chunk) {
        const long long nb = std::min(chunk, (long long)B - b0);
        if ((rc = grow(ctx, h->dParent, h->parentCap, (size_t)nb * N * 4)) != MTT_OK)
            return rc;
        MTT_CUDA(ctx, cudaMemcpyAsync(h->dParent, parent + b0 * N, (size_t)nb * N * 4,
                                      cudaMemcpyHostToDevice, s));
        const double *dH = nullptr, *dM = nullptr;
        if (mats) {
            if ((rc = grow(ctx, h->dMatsIn, h->matsInCap, (size_t)nb * N * C * 16 * 8)) != MTT_OK)
                return rc;
            MTT_CUDA(ctx, cudaMemcpyAsync(h->dMatsIn, mats + (size_t)b0 * N * C * 16,
                                          (size_t)nb * N * C * 16 * 8, cudaMemcpyHostToDevice, s));
            dM = h->dMatsIn;
        } else {
            double *&hbuf = patternLogL ? h->dMatsIn : h->dHeight;
            size_t &hcap = patternLogL ? h->matsInCap : h->heightCap;
            if ((rc = grow(ctx, hbuf, hcap, (size_t)nb * N * 8)) != MTT_OK)
                return rc;
            MTT_CUDA(ctx, cudaMemcpyAsync(hbuf, height + b0 * N, (size_t)nb * N * 8,
                                          cudaMemcpyHostToDevice, s));
            dH = hbuf;
        }
        if (patternLogL)
            dPat = h->dHeight;
        const double *dTP = nullptr;
        if (kappaPerTree) {
            std::vector<double> tp((size_t)nb * 2);
            for (long long i = 0; i < nb; i++) {
                tp[2 * i] = kappaPerTree[b0 + i];
                tp[2 * i + 1] = ratePerTree ? ratePerTree[b0 + i] : h->branchRate;
            }
            if ((rc = grow(ctx, h->dTreeParams, h->treeParamsCap, (size_t)nb * 16)) != MTT_OK)
                return rc;
            MTT_CUDA(ctx, cudaMemcpy(h->dTreeParams, tp.data(), (size_t)nb * 16,
                                     cudaMemcpyHostToDevice));
            dTP = h->dTreeParams;
        }
        const double *dNR = nullptr;
        if (nodeRates) {                 // relaxed clock: one rate per branch
            if ((rc = grow(ctx, h->dNodeRates, h->nodeRatesCap, (size_t)nb * N * 8)) != MTT_OK)
                return rc;
            MTT_CUDA(ctx, cudaMemcpyAsync(h->dNodeRates, nodeRates + b0 * N, (size_t)nb * N * 8,
                                          cudaMemcpyHostToDevice, s));
            dNR = h->dNodeRates;
        }
        rc = peel_chunk(h, nb, h->dParent, dH, dM, h->dLogL + b0, dPat, dTP,
                        dParts ? dParts + b0 : nullptr, (long long)B, dNR);
        if (rc != MTT_OK)
            return rc;
        // the upload buffers are reused by the next chunk
        MTT_CUDA(ctx, cudaStreamSynchronize(s));
    }
    if (logL)
        MTT_CUDA(ctx, cudaMemcpyAsync(logL, h->dLogL, (size_t)B * 8, cudaMemcpyDeviceToHost, s));
    if (patternLogL)
        MTT_CUDA(ctx, cudaMemcpyAsync(patternLogL, dPat, (size_t)h->nPatterns * 8,
                                      cudaMemcpyDeviceToHost, s));
    if (partsHost)
        MTT_CUDA(ctx, cudaMemcpyAsync(partsHost, dParts, (size_t)B * 24, cudaMemcpyDeviceToHost,
                                      s));
    MTT_CUDA(ctx, cudaStreamSynchronize(s));
    return MTT_OK;
}

int mtt_peel_logl_parts_batch(mtt_peel *h, int64_t B, const int32_t *parent, const double *height,
                              double *parts)
{
    int rc = check_ready(h, true);
    if (rc != MTT_OK)
        return rc;
    if (B < 0 || !parent || !height || !parts)
        return mtt_fail(h->ctx, MTT_ERR_INVALID_ARG, "mtt_peel_logl_parts_batch: bad arguments");
    if (B == 0)
        return MTT_OK;
    return peel_host(h, B, parent, height, nullptr, nullptr, nullptr, nullptr, nullptr, parts);
}

int mtt_peel_logl_batch(mtt_peel *h, int64_t B, const int32_t *parent, const double *height,
                        double *logL)
{
    int rc = check_ready(h, true);
    if (rc != MTT_OK)
        return rc;
    if (B < 0 || !parent || !height || !logL)
        return mtt_fail(h->ctx, MTT_ERR_INVALID_ARG, "mtt_peel_logl_batch: bad arguments");
    if (B == 0)
        return MTT_OK;
    return peel_host(h, B, parent, height, nullptr, logL, nullptr);
}

int mtt_peel_logl_batch_params(mtt_peel *h, int64_t B, const int32_t *parent,
                               const double *height, const double *kappa,
                               const double *branchRate, double *logL)
{
    int rc = check_ready(h, false);
    if (rc != MTT_OK)
        return rc;
    if (B < 0 || !parent || !height || !kappa || !logL)
        return mtt_fail(h->ctx, MTT_ERR_INVALID_ARG, "mtt_peel_logl_batch_params: bad arguments");
    if (B == 0)
        return MTT_OK;
    return peel_host(h, B, parent, height, nullptr, logL, nullptr, kappa, branchRate);
}

int mtt_peel_logl_batch_rates(mtt_peel *h, int64_t B, const int32_t *parent, const double *height,
                              const double *branchRates, double *logL)
{
    int rc = check_ready(h, true);
    if (rc != MTT_OK)
        return rc;
    if (B < 0 || !parent || !height || !branchRates || !logL)
        return mtt_fail(h->ctx, MTT_ERR_INVALID_ARG, "mtt_peel_logl_batch_rates: bad arguments");
    if (B == 0)
        return MTT_OK;
    return peel_host(h, B, parent, height, nullptr, logL, nullptr, nullptr, nullptr, nullptr,
                     branchRates);
}

int mtt_peel_logl_batch_mats(mtt_peel *h, int64_t B, const int32_t *parent, const double *mats,
                             double *logL)
{
    int rc = check_ready(h, false);
    if (rc != MTT_OK)
        return rc;
    if (B < 0 || !parent || !mats || !logL)
        return mtt_fail(h->ctx, MTT_ERR_INVALID_ARG, "mtt_peel_logl_batch_mats: bad arguments");
    if (B == 0)
        return MTT_OK;
    return peel_host(h, B, parent, nullptr, mats, logL, nullptr);
}

int mtt_peel_pattern_logl(mtt_peel *h, const int32_t *parent, const double *height,
                          double *patternLogL)
{
    int rc = check_ready(h, true);
    if (rc != MTT_OK)
        return rc;
    if (!parent || !height || !patternLogL)
        return mtt_fail(h->ctx, MTT_ERR_INVALID_ARG, "mtt_peel_pattern_logl: bad arguments");
    return peel_host(h, 1, parent, height, nullptr, nullptr, patternLogL);
}

}  // extern "C"
