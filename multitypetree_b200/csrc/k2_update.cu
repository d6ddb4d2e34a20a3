// k2_update.cu -- piece 2, resident chains: incremental (dirty-node) Felsenstein peeling.
//
// BEAST's TreeLikelihood [EXT] keeps the partials of every internal node between MCMC steps
// (BeerLikelihoodCore: two banks per node, currentPartialsIndex / storedPartialsIndex) and, after
// a proposal, recomputes only the nodes above a change -- the branch whose length changed, its
// parent, .. the root (TreeLikelihood.traverse: node.isDirty()); store() / restore() flip the
// bank indices.  k2_peel.cu re-evaluates whole candidate trees without ever materialising a
// partial; this file is the other regime: K chains whose partials stay RESIDENT in HBM
//     store[chain][bank][internal node][4 C state rows + 1 log-scale row][pStride]   (fp64)
// and a proposal per chain that touches a few nodes.
//   host planner     per chain: the proposal's tree is compared with the chain's accepted tree
//                    (parent, height, clock rate of every node -- or only of the nodes the caller
//                    flags dirty -- plus kappa and the strict-clock rate), applied in place under
//                    an undo journal, and the nodes to recompute are listed in post-order: the
//                    union of the paths from the changes to the root.  O(changes x depth).
//   k2_peel_update   the one kernel of an update; grid = (pattern tiles, chains with a list, the
//                    longest lists first), a thread owns one pattern column.  A CTA walks its
//                    chain's list in stages of 16 entries: the entries arrive by asynchronous
//                    copies one stage ahead, the transition matrices of their child branches are
//                    evaluated into shared memory from the entries (the closed form of
//                    k2_prepare, same argument order: bit-identical matrices).  The child that is
//                    the previous entry's result stays in registers (on a path to the root that
//                    is every entry but the first); the other child is read from its resident
//                    vector -- its lines pulled into L2 for the whole stage up front, the
//                    registers filled one entry ahead -- or from the tip states; every result is
//                    written to the bank the accepted state does not use.  Per-node rescaling as
//                    in k2_peel, the log scale factor kept CUMULATIVE per subtree (row 4 C), so
//                    the root needs no sum over nodes.  The chain's last tile to finish adds the
//                    tile sums up in tile order (what k2_finalize does).
//                    HBM-bound: (32 C + 8) bytes read + as many written per pattern and entry.
//   accept / reject  host-side: the proposal's banks, tree and logL become the accepted ones, or
//                    are forgotten.  No device work.
#include "k2_peel.cuh"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace {

struct UpdOp {          // 80 bytes
    int v;              // node recomputed (>= nTips)
    int c1, c2;         // children (node numbers); c1 is the previous entry's node when bit 0 is set
    int flags;          // bit 0: child 1 = previous result; bits 8..: chain
    double dh1, dh2;    // height(v) - height(child)
    double br1, br2;    // clock rate of the child branches
    // offsets (in doubles, row 0, pattern 0) into the resident store: the vector written -- in
    // the bank the accepted state does not use -- and the vectors of internal children
    long long dstOff, c1Off, c2Off;
    long long pad;
};
static_assert(sizeof(UpdOp) == 80, "UpdOp is staged by 16-byte copies");
constexpr int UF_PREV = 1;

// Host-side state of one chain: the ACCEPTED tree (with child lists and the bank every internal
// node's accepted partials sit in), modified in place by a proposal under an undo journal, so a
// step costs O(changed nodes x depth), not O(nodes).
struct Chain {
    // everything the planner asks about a node sits in one 40-byte record: a step of a walk to
    // the root or an entry of the list touches one or two cache lines, not eight arrays
    struct Node {
        int parent = -2;                // -2: no tree yet
        int left = -1, right = -1;
        int mark = 0, walk = 0;         // epoch / walk that reached the node (plan_chain)
        int bank = 0;                   // bank of the accepted partials (internal nodes)
        int seen = 0;                   // epoch of the sparse list that named the node
        double height = 0.0, rate = 0.0;
    };
    std::vector<Node> nd;
    int root = -1, rootCount = 0, epoch = 0, walkId = 0, seenEpoch = 0;
    double kappa = 0.0, clock = 0.0;
    double logL[4] = {0.0, 0.0, 0.0, 0.0};      // picked evaluation, unscaled sum, rescaled sum, flag
    bool valid = false, pending = false;
    // journal of the pending proposal
    struct NodeUndo { int node, parent; double height, rate; };
    struct KidsUndo { int node, left, right; };
    std::vector<NodeUndo> undoNodes;
    std::vector<KidsUndo> undoKids;
    std::vector<int> flipped, changed, moved;
    int oldRoot = -1, oldRootCount = 0;
    double oldKappa = 0.0, oldClock = 0.0, oldLogL[4] = {0.0, 0.0, 0.0, 0.0};
    bool oldValid = false;

    void begin()
    {
        undoNodes.clear();
        undoKids.clear();
        flipped.clear();
        oldRoot = root;
        oldRootCount = rootCount;
        oldKappa = kappa;
        oldClock = clock;
        memcpy(oldLogL, logL, sizeof logL);
        oldValid = valid;
        pending = true;
    }
    void undo()
    {
        for (size_t i = undoKids.size(); i-- > 0;) {
            nd[undoKids[i].node].left = undoKids[i].left;
            nd[undoKids[i].node].right = undoKids[i].right;
        }
        for (size_t i = undoNodes.size(); i-- > 0;) {
            const NodeUndo &u = undoNodes[i];
            nd[u.node].parent = u.parent;
            nd[u.node].height = u.height;
            nd[u.node].rate = u.rate;
        }
        for (int v : flipped)
            nd[v].bank ^= 1;
        root = oldRoot;
        rootCount = oldRootCount;
        kappa = oldKappa;
        clock = oldClock;
        memcpy(logL, oldLogL, sizeof logL);
        valid = oldValid;
        pending = false;
        undoNodes.clear();
        undoKids.clear();
        flipped.clear();
    }
    void commit()
    {
        pending = false;
        undoNodes.clear();
        undoKids.clear();
        flipped.clear();
    }
};

}  // namespace

struct mtt_peel_chains {
    mtt_peel *h = nullptr;
    int K = 0;
    long long version = -1;             // h->version the accepted states were computed under
    double *dStore = nullptr;           // resident partials
    double *dPart = nullptr; size_t partCap = 0;     // tile sums [chain][tile][3]
    int *dDone = nullptr;               // tiles finished per chain (self-resetting)
    std::vector<Chain> chains;
    // planner scratch
    std::vector<UpdOp> ops;
    std::vector<int> opOff, stack, order;
    long long lastOps = 0;
    int unflagged = -1;                 // node found by the verification below
    bool needFull = false;              // a sparse update met a chain without an accepted state
    std::vector<int32_t> spParent;      // sparse updates: the listed nodes' values by node number
    std::vector<double> spHeight, spRate;
    // integration aid (mtt_peel_chains_set_verify): with a dirty mask, compare ALL nodes anyway
    // and fail the call when a node differs without being flagged
    bool verify = false;
    // optional device timing of the launches of an update (mtt_peel_chains_set_timing)
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float lastMs = 0.0f;
};

namespace {

#ifndef MTT_K2U_THREADS
#define MTT_K2U_THREADS 128
#endif
constexpr int UPD_THREADS = MTT_K2U_THREADS;

__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit()
{
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// entries per stage: list entries + their matrices go through shared memory, double-buffered
__host__ __device__ constexpr int upd_stage_entries(int C)
{
    return C <= 2 ? 16 : (C <= 4 ? 8 : 4);
}

template <int C>
#ifndef MTT_K2U_MINB
#define MTT_K2U_MINB 6
#endif
__global__ void __launch_bounds__(UPD_THREADS, C == 1 ? MTT_K2U_MINB : 1)
k2_peel_update(int n, long long P, long long pStride, const uint8_t *states,
               const double *weights, const uint8_t *constMask, const double *site, double pInv,
               const UpdOp *ops, const int *opOff, const int *chainOrder,
               const double *kappaPerChain, double *store, double *partial, int nTiles,
               int *tilesDone, double *logL, long long nChains)
{
    constexpr int CH = upd_stage_entries(C);
    constexpr int OPB = (int)sizeof(UpdOp), MATB = 2 * C * 16 * 8, STAGE = CH * (OPB + MATB);
    __shared__ __align__(16) unsigned char stage[2][STAGE];
    __shared__ double red[3 * (UPD_THREADS / 32)];
    __shared__ int sLast;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // chains in order of decreasing list length: the long walks start first (only chains with a
    // non-empty list are launched)
    const int tile = blockIdx.x, chain = chainOrder[blockIdx.y];
    const long long p = (long long)tile * UPD_THREADS + tid;        // < pStride
    const int k0 = opOff[chain], k1 = opOff[chain + 1];
    double *out = partial + ((size_t)chain * nTiles + tile) * 3;
    constexpr int ROWS = 4 * C + 1;
    const double kappa = kappaPerChain[chain];
    double *const mine = store + p;     // + vector offset + row * pStride

    // a stage: CH list entries (asynchronous copies, one stage ahead) and the transition
    // matrices of their child branches, evaluated here from the entries -- the closed form of
    // k2_prepare with the same argument order, so the matrices are bit-identical to the full
    // evaluation's: [entry][child][category][16], internal child row-major, tip child transposed
    auto issue = [&](int j) {
        const int e0 = k0 + j * CH, cnt = min(CH, k1 - e0);
        unsigned char *dst = stage[j & 1];
        const unsigned char *srcO = reinterpret_cast<const unsigned char *>(ops + e0);
        for (int o = tid * 16; o < cnt * OPB; o += UPD_THREADS * 16)
            cp_async16(dst + o, srcO + o);
        cp_async_commit();
    };
    auto matrices = [&](int j) {
        const int cnt = min(CH, k1 - k0 - j * CH);
        const UpdOp *sOps = reinterpret_cast<const UpdOp *>(stage[j & 1]);
        double *sMats = reinterpret_cast<double *>(stage[j & 1] + CH * OPB);
        const double *catRates = site, *freqs = site + 2 * C;
        for (int idx = tid; idx < cnt * 2 * C; idx += UPD_THREADS) {
            const int e = idx / (2 * C), jj = (idx / C) & 1, c = idx % C;
            const UpdOp &op = sOps[e];
            const int child = jj ? op.c2 : op.c1;
            const double jointBranchRate = catRates[c] * (jj ? op.br2 : op.br1);
            const double distance = (jj ? op.dh2 : op.dh1) * jointBranchRate;
            double T[16];
            hky_matrix(kappa, freqs, distance, T);
            double *M = sMats + ((size_t)(e * 2 + jj) * C + c) * 16;
            const bool tip = child < n;
#pragma unroll
            for (int q = 0; q < 16; q++)
                M[q] = tip ? T[(q & 3) * 4 + (q >> 2)] : T[q];
        }
    };
    const int nStages = (k1 - k0 + CH - 1) / CH;
    issue(0);
    cp_async_wait<0>();
    __syncthreads();

    double prev[C][1][4], prevScale = 0.0;
    // child 2 of the entry about to run, fetched one entry ahead
    double nx[C][1][4], nScale = 0.0;
    int nTip[1] = {4};
    auto fetch2 = [&](const UpdOp &op) {
        if (op.c2 < n) {
            nTip[0] = __ldg(states + (size_t)op.c2 * pStride + p);
        } else {
            const double *src = mine + op.c2Off;
#pragma unroll
            for (int r = 0; r < 4 * C; r++, src += pStride)
                nx[r >> 2][0][r & 3] = __ldcg(src);
            nScale = __ldcg(src);
        }
    };
#pragma unroll 1
    for (int j = 0; j < nStages; j++) {
        // (this stage's entries have landed; the other buffer was released by the barrier at the
        // end of the previous round)
        if (j + 1 < nStages)
            issue(j + 1);
        const UpdOp *sOps = reinterpret_cast<const UpdOp *>(stage[j & 1]);
        const double *sMats = reinterpret_cast<const double *>(stage[j & 1] + CH * OPB);
        const int cnt = min(CH, k1 - k0 - j * CH);
        // The resident children of a path do not depend on the entries before them: pull the
        // lines of the whole stage into L2 now (one lane per 128-byte line), so that the
        // register fetch one entry ahead is an L2 hit and the walk is not a chain of HBM
        // latencies.
        if ((lane & 15) == 0) {
            for (int e = 1; e < cnt; e++) {
                if (sOps[e].c2 < n)
                    continue;
                const double *src = mine + sOps[e].c2Off;
#pragma unroll
                for (int r = 0; r < ROWS; r++, src += pStride)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(src));
            }
        }
        fetch2(sOps[0]);
        // ... and while those are in flight, the stage's transition matrices
        matrices(j);
        __syncthreads();
#pragma unroll 1
        for (int e = 0; e < cnt; e++) {
            const UpdOp &op = sOps[e];
            const int c1 = op.c1, c2 = op.c2, flags = op.flags;
            const double *M1 = sMats + (size_t)e * 2 * C * 16, *M2 = M1 + C * 16;
            // child 1: the previous result, a tip, or (first entry of a path) a resident vector
            int t1[1] = {4};
            double ls = 0.0;
            if (flags & UF_PREV) {
                ls = prevScale;
            } else if (c1 < n) {
                t1[0] = __ldg(states + (size_t)c1 * pStride + p);
            } else {
                const double *src = mine + op.c1Off;
#pragma unroll
                for (int r = 0; r < 4 * C; r++, src += pStride)
                    prev[r >> 2][0][r & 3] = __ldcg(src);
                ls = __ldcg(src);
            }
            // child 2 (already here); start on the next entry's
            double x2[C][1][4];
            int t2[1] = {nTip[0]};
#pragma unroll
            for (int c = 0; c < C; c++)
#pragma unroll
                for (int i = 0; i < 4; i++)
                    x2[c][0][i] = nx[c][0][i];
            if (c2 >= n)
                ls += nScale;
            double *dst = mine + op.dstOff;
            if (e + 1 < cnt)
                fetch2(sOps[e + 1]);
            int mxhi = 0;
#pragma unroll
            for (int c = 0; c < C; c++) {
                double a[1][4], bb[1][4];
                child_term<1, 2>(c1 < n, M1 + c * 16, prev[c], t1, a);
                child_term<1, 2>(c2 < n, M2 + c * 16, x2[c], t2, bb);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    prev[c][0][i] = a[0][i] * bb[0][i];
                    mxhi = max(mxhi, __double2hiint(prev[c][0][i]));
                }
            }
            // BeerLikelihoodCore.scalePartials [EXT], as in k2_peel
            if (mxhi <= __double2hiint(SCALING_THRESHOLD)) {
                double mx = 0.0;
#pragma unroll
                for (int c = 0; c < C; c++)
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        mx = fmax(mx, prev[c][0][i]);
                if (mx < SCALING_THRESHOLD) {
#pragma unroll
                    for (int c = 0; c < C; c++)
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            prev[c][0][i] /= mx;
                    ls += log(mx);
                }
            }
            prevScale = ls;
#pragma unroll
            for (int r = 0; r < 4 * C; r++, dst += pStride)
                *dst = prev[r >> 2][0][r & 3];
            *dst = ls;
        }
        if (j + 1 < nStages)
            cp_async_wait<0>();
        __syncthreads();                // done with this stage; the next one's entries have landed
    }

    // the last entry is the root
    const double *catProps = site + C, *freqs = site + 2 * C;
    double root[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double r = prev[0][0][i] * catProps[0];
#pragma unroll
        for (int c = 1; c < C; c++)
            r += prev[c][0][i] * catProps[c];
        root[i] = r;
    }
    double llU, llQ;
    bool under;
    root_pattern_logl(root, prevScale, pInv, constMask, p, freqs, llU, llQ, under);
    double contribU = 0.0, contribQ = 0.0;
    bool anyUnder = false;
    if (p < P) {
        contribU = llU * weights[p];
        contribQ = llQ * weights[p];
        anyUnder = under;
    }
    contribU = warp_sum_down(contribU);
    contribQ = warp_sum_down(contribQ);
    anyUnder = __any_sync(FULL_MASK, anyUnder);
    if (lane == 0) {
        red[warp] = contribU;
        red[UPD_THREADS / 32 + warp] = contribQ;
        red[2 * (UPD_THREADS / 32) + warp] = anyUnder ? 1.0 : 0.0;
    }
    __syncthreads();
    if (tid == 0) {
        double tU = 0.0, tQ = 0.0, f = 0.0;
#pragma unroll
        for (int w = 0; w < UPD_THREADS / 32; w++) {
            tU += red[w];
            tQ += red[UPD_THREADS / 32 + w];
            f += red[2 * (UPD_THREADS / 32) + w];
        }
        out[0] = tU;
        out[1] = tQ;
        out[2] = f;
        // the chain's last tile to finish adds the tile sums up, in tile order (k2_finalize)
        __threadfence();
        sLast = atomicAdd(&tilesDone[chain], 1) == nTiles - 1;
    }
    __syncthreads();
    if (sLast && warp == 0) {
        __threadfence();
        const double *in = partial + (size_t)chain * nTiles * 3;
        double sU = 0.0, sQ = 0.0, fl = 0.0;
        for (int t0 = 0; t0 < nTiles; t0 += 32) {
            const int t = t0 + lane;
            const double u = t < nTiles ? __ldcg(in + (size_t)t * 3) : 0.0;
            const double q = t < nTiles ? __ldcg(in + (size_t)t * 3 + 1) : 0.0;
            const double g = t < nTiles ? __ldcg(in + (size_t)t * 3 + 2) : 0.0;
            // lane order = tile order
            for (int l = 0; l < 32 && t0 + l < nTiles; l++) {
                sU += __shfl_sync(FULL_MASK, u, l);
                sQ += __shfl_sync(FULL_MASK, q, l);
                fl += __shfl_sync(FULL_MASK, g, l);
            }
        }
        if (lane == 0) {
            // logL[4][nChains]: the evaluation BEAST would return, and the parts a pattern shard
            // hands to the all-reduce (unscaled sum, rescaled sum, tiles that underflowed)
            logL[chain] = (fl > 0.0) ? sQ : sU;
            logL[nChains + chain] = sU;
            logL[2 * nChains + chain] = sQ;
            logL[3 * nChains + chain] = fl;
            tilesDone[chain] = 0;       // ready for the next update
        }
    }
}

template <int C>
int launch_update(mtt_peel_chains *s, const UpdOp *dOps, const int *dOpOff, const int *dOrder,
                  int nActive, int nTiles, const double *dKappa, double *dLogL)
{
    mtt_peel *h = s->h;
    mtt_ctx *ctx = h->ctx;
    dim3 grid((unsigned)nTiles, (unsigned)nActive);
    k2_peel_update<C><<<grid, UPD_THREADS, 0, ctx->stream>>>(
        h->nTips, h->nPatterns, h->pStride, h->dStates, h->dWeights, h->dConst, h->dSite, h->pInv,
        dOps, dOpOff, dOrder, dKappa, s->dStore, s->dPart, nTiles, s->dDone, dLogL, s->K);
    MTT_CHECK_LAUNCH(ctx);
    return MTT_OK;
}

// Apply the proposal (parent, height, rate) of one chain to its state and append, in post-order,
// the nodes that must be recomputed: an internal node is listed when the branch of a child
// changed (the child's height, clock rate or parent, or the node's own height) or a child is
// listed itself -- the union of the paths from the changes to the root.  `dirty` (optional, N
// flags) names the nodes that MAY differ from the accepted tree; without it every node is
// compared; `list` (optional, instead of dirty) is the same thing as a list of node numbers --
// parent / height / rate are then only read at the listed nodes.  Returns false for a malformed
// tree (the state is rolled back).
bool plan_chain(mtt_peel_chains *s, int k, const int32_t *parent, const double *height,
                const double *rate, double clock, double kappa, const uint8_t *dirty,
                const int32_t *list = nullptr, int nList = 0)
{
    const int n = s->h->nTips, N = 2 * n - 1;
    Chain &c = s->chains[k];
    const bool all = !c.valid || kappa != c.kappa || (!rate && clock != c.clock);
    if (all && list) {
        s->needFull = true;             // a list cannot describe a whole tree
        return false;
    }
    c.begin();
    c.kappa = kappa;
    c.clock = clock;
    std::vector<int> &changed = c.changed;
    changed.clear();
    auto differs = [&](int i) {
        return parent[i] != c.nd[i].parent || height[i] != c.nd[i].height ||
               (rate ? rate[i] : clock) != c.nd[i].rate;
    };
    if (all) {
        for (int i = 0; i < N; i++)
            changed.push_back(i);
    } else if (list) {
        if (c.seenEpoch == 0x7fffffff) {
            for (Chain::Node &x : c.nd)
                x.seen = 0;
            c.seenEpoch = 0;
        }
        const int tag = ++c.seenEpoch;
        for (int j = 0; j < nList; j++) {
            const int i = list[j];
            if (c.nd[i].seen == tag)
                continue;               // named twice
            c.nd[i].seen = tag;
            if (differs(i))
                changed.push_back(i);
        }
    } else if (dirty) {
        if (s->verify)
            for (int i = 0; i < N; i++)
                if (!dirty[i] && differs(i)) {
                    s->unflagged = i;       // reported by the caller of plan_chain
                    c.undo();
                    return false;
                }
        // a handful of flags in N bytes: skip eight clean nodes at a time
        int i = 0;
        for (; i + 8 <= N; i += 8) {
            uint64_t w;
            memcpy(&w, dirty + i, 8);
            if (!w)
                continue;
            for (int j = i; j < i + 8; j++)
                if (dirty[j] && differs(j))
                    changed.push_back(j);
        }
        for (; i < N; i++)
            if (dirty[i] && differs(i))
                changed.push_back(i);
    } else {
        for (int i = 0; i < N; i++)
            if (differs(i))
                changed.push_back(i);
    }
    if (changed.empty())
        return true;
    auto fail = [&]() {
        c.undo();
        return false;
    };
    auto journalKids = [&](int v) { c.undoKids.push_back({v, c.nd[v].left, c.nd[v].right}); };
    // detach the nodes whose parent changes, then attach them to their new parents
    std::vector<int> &touched = s->stack;
    touched.clear();
    for (int i : changed) {
        const int pp = parent[i];
        if (pp >= N || pp == i || (pp >= 0 && pp < n))
            return fail();
        c.undoNodes.push_back({i, c.nd[i].parent, c.nd[i].height, c.nd[i].rate});
        if (pp != c.nd[i].parent) {
            const int old = c.nd[i].parent;
            if (old >= n && old < N) {
                journalKids(old);
                if (c.nd[old].left == i)
                    c.nd[old].left = -1;
                else if (c.nd[old].right == i)
                    c.nd[old].right = -1;
                touched.push_back(old);
            } else if (old == -1) {
                c.rootCount--;
            }
        }
    }
    // (an internal node whose own height moved has two changed child branches: noted in `moved`)
    std::vector<int> &moved = c.moved;
    moved.clear();
    for (int i : changed) {
        const int pp = parent[i];
        if (i >= n && (all || height[i] != c.nd[i].height))
            moved.push_back(i);
        if (pp != c.nd[i].parent) {
            if (pp >= 0) {
                journalKids(pp);
                if (c.nd[pp].left < 0)
                    c.nd[pp].left = i;
                else if (c.nd[pp].right < 0)
                    c.nd[pp].right = i;
                else
                    return fail();                  // a third child
                touched.push_back(pp);
            } else {
                c.rootCount++;
                c.root = i;
            }
            c.nd[i].parent = pp;
        }
        c.nd[i].height = height[i];
        c.nd[i].rate = rate ? rate[i] : clock;
    }
    if (c.rootCount != 1 || c.root < n)
        return fail();
    for (int v : touched)
        if (c.nd[v].left < 0 || c.nd[v].right < 0)
            return fail();                          // an internal node left with one child
    // walks from the changes up to the root (or to a node already marked in this call); meeting
    // a node marked by the SAME walk means a cycle
    if (c.epoch == 0x7fffffff) {
        for (Chain::Node &x : c.nd)
            x.mark = 0;
        c.epoch = 0;
    }
    c.epoch++;
    c.walkId = 0;
    auto walkUp = [&](int u) {
        const int id = ++c.walkId;
        int steps = 0;
        while (u >= 0) {
            if (c.nd[u].mark == c.epoch) {
                if (c.nd[u].walk == id)
                    return false;
                return true;
            }
            c.nd[u].mark = c.epoch;
            c.nd[u].walk = id;
            u = c.nd[u].parent;
            if (++steps > N)
                return false;
        }
        return true;
    };
    for (int i : changed)
        if (c.nd[i].parent >= 0 && !walkUp(c.nd[i].parent))
            return fail();
    for (int i : moved)
        if (!walkUp(i))
            return fail();
    if (c.nd[c.root].mark != c.epoch)
        return true;                                // e.g. only the root's branch rate changed
    // post-order over the marked nodes; a marked child is visited last so that it is the
    // previous entry of its parent (two marked children: the other one is read back from HBM).
    // Emitted in reverse post-order, reversed afterwards.
    std::vector<int> &stack = s->stack;
    stack.clear();
    stack.push_back(c.root);
    const size_t first = s->ops.size();
    while (!stack.empty()) {
        const int v = stack.back();
        stack.pop_back();
        const int l = c.nd[v].left, r = c.nd[v].right;
        const bool lu = l >= n && c.nd[l].mark == c.epoch, ru = r >= n && c.nd[r].mark == c.epoch;
        UpdOp op;
        memset(&op, 0, sizeof(op));
        op.v = v;
        const int c1 = ru ? r : l, c2 = ru ? l : r;
        op.c1 = c1;
        op.c2 = c2;
        op.flags = (k << 8) | ((lu || ru) ? UF_PREV : 0);
        c.nd[v].bank ^= 1;                          // the bank the accepted state does not use
        c.flipped.push_back(v);
        op.dh1 = c.nd[v].height - c.nd[c1].height;
        op.dh2 = c.nd[v].height - c.nd[c2].height;
        op.br1 = c.nd[c1].rate;
        op.br2 = c.nd[c2].rate;
        s->ops.push_back(op);
        if (ru && lu)
            stack.push_back(c2);
        if (lu || ru)
            stack.push_back(c1);
    }
    std::reverse(s->ops.begin() + first, s->ops.end());
    // where the vectors live, known once every listed node has its new bank
    const long long rows = 4 * s->h->nCat + 1, nInt = n - 1;
    auto vecOff = [&](int node) {
        return (((long long)k * 2 + c.nd[node].bank) * nInt + (node - n)) * rows * s->h->pStride;
    };
    for (size_t e = first; e < s->ops.size(); e++) {
        UpdOp &op = s->ops[e];
        op.dstOff = vecOff(op.v);
        op.c1Off = op.c1 >= n ? vecOff(op.c1) : 0;
        op.c2Off = op.c2 >= n ? vecOff(op.c2) : 0;
    }
    return true;
}

}  // namespace

// free the device side; called by mtt_peel_destroy for chains that outlive their handle (they
// answer MTT_ERR_STATE from then on)
void k2_chains_release(mtt_peel_chains *s)
{
    if (!s->h)
        return;
    cudaSetDevice(s->h->ctx->device);
    cudaStreamSynchronize(s->h->ctx->stream);
    for (void *p : {(void *)s->dStore, (void *)s->dPart, (void *)s->dDone})
        if (p)
            cudaFree(p);
    if (s->ev0)
        cudaEventDestroy(s->ev0);
    if (s->ev1)
        cudaEventDestroy(s->ev1);
    s->dStore = s->dPart = nullptr;
    s->dDone = nullptr;
    s->ev0 = s->ev1 = nullptr;
    s->h = nullptr;
}

extern "C" {

int mtt_peel_chains_create(mtt_peel *h, int nChains, mtt_peel_chains **out)
{
    if (!h || !out)
        return MTT_ERR_INVALID_ARG;
    *out = nullptr;
    mtt_ctx *ctx = h->ctx;
    if (nChains < 1 || nChains >= (1 << 22))
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "mtt_peel_chains_create: nChains=%d", nChains);
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)nChains * 2 * (size_t)(h->nTips - 1) * (4 * h->nCat + 1) *
                         (size_t)h->pStride * 8;
    size_t freeB = 0, totalB = 0;
    MTT_CUDA(ctx, cudaMemGetInfo(&freeB, &totalB));
    if (bytes + ((size_t)1 << 30) > freeB)
        return mtt_fail(ctx, MTT_ERR_TOO_LARGE,
                        "mtt_peel_chains_create: %d chains need %.1f GB of resident partials, "
                        "%.1f GB free", nChains, bytes / 1e9, freeB / 1e9);
    mtt_peel_chains *s = new mtt_peel_chains();
    s->h = h;
    s->K = nChains;
    cudaError_t e = cudaMalloc((void **)&s->dStore, bytes + 256);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&s->dDone, (size_t)nChains * sizeof(int));
    if (e == cudaSuccess)
        e = cudaMemset(s->dDone, 0, (size_t)nChains * sizeof(int));
    if (e != cudaSuccess) {
        if (s->dStore)
            cudaFree(s->dStore);
        delete s;
        return mtt_fail(ctx, MTT_ERR_CUDA, "mtt_peel_chains_create: cudaMalloc(%zu) failed: %s",
                        bytes, cudaGetErrorString(e));
    }
    const int N = 2 * h->nTips - 1;
    s->chains.resize(nChains);
    for (Chain &c : s->chains)
        c.nd.assign(N, Chain::Node());
    h->chains.push_back(s);
    *out = s;
    return MTT_OK;
}

int mtt_peel_chains_destroy(mtt_peel_chains *s)
{
    if (!s)
        return MTT_OK;
    if (s->h) {
        std::vector<mtt_peel_chains *> &v = s->h->chains;
        v.erase(std::remove(v.begin(), v.end(), s), v.end());
        k2_chains_release(s);
    }
    delete s;
    return MTT_OK;
}

// device side of an update: the lists of all chains in one launch; got[4][K]
static int chains_launch(mtt_peel_chains *s, const std::vector<double> &kap,
                         std::vector<double> &got)
{
    mtt_peel *h = s->h;
    mtt_ctx *ctx = h->ctx;
    const int C = h->nCat, K = s->K;
    int rc;
    const int nTiles = (int)(h->pStride / UPD_THREADS);
    if ((rc = grow(ctx, s->dPart, s->partCap, (size_t)K * nTiles * 24)) != MTT_OK)
        return rc;
    Xfer x(ctx);
    UpdOp *dOps = nullptr;
    int *dOpOff = nullptr;
    double *dKappa = nullptr, *dLogL = nullptr;
    // chains with something to do, the longest lists first
    std::vector<int> &order = s->order;
    order.clear();
    for (int k = 0; k < K; k++)
        if (s->opOff[k + 1] > s->opOff[k])
            order.push_back(k);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return s->opOff[a + 1] - s->opOff[a] > s->opOff[b + 1] - s->opOff[b];
    });
    const int nActive = (int)order.size();
    int *dOrder = nullptr;
    x.in(s->ops.data(), s->ops.size(), BUF_C_OPS, &dOps);
    x.in(s->opOff.data(), s->opOff.size(), BUF_C_OPOFF, &dOpOff);
    x.in(order.data(), order.size(), BUF_C_ORDER, &dOrder);
    x.in(kap.data(), kap.size(), BUF_C_KAPPA, &dKappa);
    x.out(got.data(), got.size(), BUF_C_LOGL, &dLogL);
    x.direct_out();                         // written once, by the last tile of each chain
    if ((rc = x.prepare()) != MTT_OK || (rc = x.upload()) != MTT_OK)
        return rc;
    if (s->timing)
        MTT_CUDA(ctx, cudaEventRecord(s->ev0, ctx->stream));
    switch (C) {
    case 1: rc = launch_update<1>(s, dOps, dOpOff, dOrder, nActive, nTiles, dKappa, dLogL); break;
    case 2: rc = launch_update<2>(s, dOps, dOpOff, dOrder, nActive, nTiles, dKappa, dLogL); break;
    case 3: rc = launch_update<3>(s, dOps, dOpOff, dOrder, nActive, nTiles, dKappa, dLogL); break;
    case 4: rc = launch_update<4>(s, dOps, dOpOff, dOrder, nActive, nTiles, dKappa, dLogL); break;
    case 5: rc = launch_update<5>(s, dOps, dOpOff, dOrder, nActive, nTiles, dKappa, dLogL); break;
    case 6: rc = launch_update<6>(s, dOps, dOpOff, dOrder, nActive, nTiles, dKappa, dLogL); break;
    case 7: rc = launch_update<7>(s, dOps, dOpOff, dOrder, nActive, nTiles, dKappa, dLogL); break;
    case 8: rc = launch_update<8>(s, dOps, dOpOff, dOrder, nActive, nTiles, dKappa, dLogL); break;
    default: rc = mtt_fail(ctx, MTT_ERR_TOO_LARGE, "peeling: nCat=%d outside [1,8]", C);
    }
    if (rc != MTT_OK)
        return rc;
    if (s->timing)
        MTT_CUDA(ctx, cudaEventRecord(s->ev1, ctx->stream));
    if ((rc = x.download()) != MTT_OK)
        return rc;
    if (s->timing)
        MTT_CUDA(ctx, cudaEventElapsedTime(&s->lastMs, s->ev0, s->ev1));
    return MTT_OK;
}

// spOff / spNode (sparse form): chain k's proposal is the list of nodes
// spNode[spOff[k] .. spOff[k+1]) with parent / height / branchRates given per LIST ENTRY
static int chains_update(mtt_peel_chains *s, const int32_t *parent, const double *height,
                         const double *branchRates, const double *kappa, const double *clockRate,
                         const uint8_t *dirty, double *logL, double *parts,
                         const int64_t *spOff = nullptr, const int32_t *spNode = nullptr)
{
    if (!s || !s->h)
        return MTT_ERR_STATE;               // (the handle was destroyed)
    mtt_peel *h = s->h;
    mtt_ctx *ctx = h->ctx;
    int rc = k2_check_ready(h, kappa == nullptr);
    if (rc != MTT_OK)
        return rc;
    if (!parent || !height || (!logL && !parts))
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "mtt_peel_chains_update: bad arguments");
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    const int n = h->nTips, N = 2 * n - 1, K = s->K;
    for (Chain &c : s->chains)
        if (c.pending)
            c.undo();                      // an unanswered proposal counts as rejected
    if (s->version != h->version) {         // alignment / site model changed: nothing is reusable
        for (Chain &c : s->chains)
            c.valid = false;
        s->version = h->version;
    }
    s->ops.clear();
    s->opOff.assign(K + 1, 0);
    std::vector<double> kap(K);
    std::vector<uint8_t> bad(K, 0);
    for (int k = 0; k < K; k++) {
        const double kk = kappa ? kappa[k] : h->kappa;
        const double clock = clockRate ? clockRate[k] : h->branchRate;
        const double *rate = (branchRates && !spOff) ? branchRates + (size_t)k * N : nullptr;
        kap[k] = kk;
        const size_t before = s->ops.size();
        s->unflagged = -1;
        s->needFull = false;
        bool planned;
        if (spOff) {
            // the listed nodes' values, by node number (only those entries are read)
            s->spParent.resize(N);
            s->spHeight.resize(N);
            s->spRate.resize(N);
            const int64_t lo = spOff[k], hi = spOff[k + 1];
            bool inRange = hi >= lo;
            for (int64_t j = lo; j < hi && inRange; j++) {
                const int i = spNode[j];
                if (i < 0 || i >= N) {
                    inRange = false;
                    break;
                }
                s->spParent[i] = parent[j];
                s->spHeight[i] = height[j];
                if (branchRates)
                    s->spRate[i] = branchRates[j];
            }
            planned = inRange &&
                      plan_chain(s, k, s->spParent.data(), s->spHeight.data(),
                                 branchRates ? s->spRate.data() : nullptr, clock, kk, nullptr,
                                 spNode + lo, (int)(hi - lo));
        } else {
            planned = plan_chain(s, k, parent + (size_t)k * N, height + (size_t)k * N, rate, clock,
                                 kk, dirty ? dirty + (size_t)k * N : nullptr);
        }
        if (!planned) {
            s->ops.resize(before);
            bad[k] = 1;
            if (s->needFull) {
                for (Chain &c : s->chains)
                    if (c.pending)
                        c.undo();
                return mtt_fail(ctx, MTT_ERR_STATE,
                                "mtt_peel_chains_update_sparse: chain %d has no accepted tree "
                                "(or its kappa / clock rate changed): a full update is needed", k);
            }
            if (s->unflagged >= 0) {
                for (Chain &c : s->chains)
                    if (c.pending)
                        c.undo();
                return mtt_fail(ctx, MTT_ERR_INVALID_ARG,
                                "mtt_peel_chains_update: chain %d, node %d differs from the "
                                "accepted tree but is not flagged dirty", k, s->unflagged);
            }
        }
        s->opOff[k + 1] = (int)s->ops.size();
    }
    const long long nEntries = (long long)s->ops.size();
    s->lastOps = nEntries;
    s->lastMs = 0.0f;
    if (nEntries > 0) {
        std::vector<double> got((size_t)4 * K);
        if ((rc = chains_launch(s, kap, got)) != MTT_OK) {
            for (Chain &c : s->chains)
                if (c.pending)
                    c.undo();               // no half-evaluated proposals: every chain as accepted
            return rc;
        }
        for (int k = 0; k < K; k++)
            if (s->opOff[k + 1] > s->opOff[k])
                for (int q = 0; q < 4; q++)
                    s->chains[k].logL[q] = got[(size_t)q * K + k];
    }
    for (int k = 0; k < K; k++) {
        Chain &c = s->chains[k];
        if (!bad[k])
            c.valid = true;                 // (of the proposal; undo() restores the old flag)
        // malformed tree: NaN, loudly, and the chain keeps its state; otherwise the proposal's
        // values -- the accepted ones when nothing had to be recomputed
        if (logL)
            logL[k] = bad[k] ? NAN : c.logL[0];
        if (parts)
            for (int q = 0; q < 3; q++)
                parts[(size_t)q * K + k] = bad[k] ? NAN : c.logL[q + 1];
    }
    return MTT_OK;
}

int mtt_peel_chains_update(mtt_peel_chains *s, const int32_t *parent, const double *height,
                           const double *branchRates, const double *kappa,
                           const double *clockRate, const uint8_t *dirty, double *logL)
{
    return chains_update(s, parent, height, branchRates, kappa, clockRate, dirty, logL, nullptr);
}

int mtt_peel_chains_update_sparse(mtt_peel_chains *s, const int64_t *off, const int32_t *node,
                                  const int32_t *parent, const double *height,
                                  const double *branchRates, const double *kappa,
                                  const double *clockRate, double *logL)
{
    if (s && s->h && (!off || (!node && off[s->K] > 0)))
        return mtt_fail(s->h->ctx, MTT_ERR_INVALID_ARG, "mtt_peel_chains_update_sparse: bad arguments");
    static const int32_t none32 = 0;
    static const double none64 = 0.0;
    // (an empty proposal has no arrays to point at)
    return chains_update(s, parent ? parent : &none32, height ? height : &none64, branchRates,
                         kappa, clockRate, nullptr, logL, nullptr, off, node ? node : &none32);
}

int mtt_peel_chains_update_parts(mtt_peel_chains *s, const int32_t *parent, const double *height,
                                 const double *branchRates, const double *kappa,
                                 const double *clockRate, const uint8_t *dirty, double *parts)
{
    return chains_update(s, parent, height, branchRates, kappa, clockRate, dirty, nullptr, parts);
}

int mtt_peel_chains_accept(mtt_peel_chains *s, const uint8_t *accept)
{
    if (!s || !s->h)
        return MTT_ERR_STATE;
    for (int k = 0; k < s->K; k++) {
        Chain &c = s->chains[k];
        if (!c.pending)
            continue;
        if (!accept || accept[k])
            c.commit();
        else
            c.undo();
    }
    return MTT_OK;
}

int64_t mtt_peel_chains_last_entries(const mtt_peel_chains *s)
{
    return s ? s->lastOps : 0;
}

int mtt_peel_chains_set_timing(mtt_peel_chains *s, int on)
{
    if (!s || !s->h)
        return MTT_ERR_STATE;
    mtt_ctx *ctx = s->h->ctx;
    if (on && !s->ev0) {
        MTT_CUDA(ctx, cudaSetDevice(ctx->device));
        MTT_CUDA(ctx, cudaEventCreate(&s->ev0));
        MTT_CUDA(ctx, cudaEventCreate(&s->ev1));
    }
    s->timing = on != 0;
    return MTT_OK;
}

int mtt_peel_chains_set_verify(mtt_peel_chains *s, int on)
{
    if (!s || !s->h)
        return MTT_ERR_STATE;
    s->verify = on != 0;
    return MTT_OK;
}

double mtt_peel_chains_last_device_ms(const mtt_peel_chains *s)
{
    return s ? (double)s->lastMs : 0.0;
}

}  // extern "C"
