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
//   k2_update_mats   transition matrices of the two child branches of every listed node (the
//                    closed form of k2_prepare, same argument order: bit-identical matrices).
//   k2_peel_update   grid = (pattern tiles, chains), a thread owns one pattern column.  A CTA
//                    walks its chain's list; the child that is the previous entry's result stays
//                    in registers (on a path to the root that is every entry but the first), the
//                    other child is read from its resident vector -- fetched one entry ahead --
//                    or from the tip states; every result is written to the bank the accepted
//                    state does not use.  Per-node rescaling as in k2_peel, the log scale factor
//                    kept CUMULATIVE per subtree (row 4 C), so the root needs no sum over nodes.
//                    HBM-bound: (32 C + 8) bytes read + as many written per pattern and entry.
//   accept / reject  host-side: the proposal's banks, tree and logL become the accepted ones, or
//                    are forgotten.  No device work.
#include "k2_peel.cuh"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace {

struct UpdOp {          // 48 bytes
    int v;              // node recomputed (>= nTips)
    int c1, c2;         // children (node numbers); c1 is the previous entry's node when bit 0 is set
    int flags;          // bit 0: child 1 = previous result; bit 1: bank written; bits 2, 3: banks of
                        // child 1, child 2 (internal, resident); bits 8..: chain
    double dh1, dh2;    // height(v) - height(child)
    double br1, br2;    // clock rate of the child branches
};
constexpr int UF_PREV = 1, UF_DST_BANK = 2, UF_C1_BANK = 4, UF_C2_BANK = 8;

// Host-side state of one chain: the ACCEPTED tree (with child lists and the bank every internal
// node's accepted partials sit in), modified in place by a proposal under an undo journal, so a
// step costs O(changed nodes x depth), not O(nodes).
struct Chain {
    std::vector<int> parent, left, right, mark, walk;
    std::vector<double> height, rate;
    std::vector<uint8_t> bank;          // per internal node
    int root = -1, rootCount = 0, epoch = 0, walkId = 0;
    double kappa = 0.0, clock = 0.0, logL = 0.0;
    bool valid = false, pending = false;
    // journal of the pending proposal
    struct NodeUndo { int node, parent; double height, rate; };
    struct KidsUndo { int node, left, right; };
    std::vector<NodeUndo> undoNodes;
    std::vector<KidsUndo> undoKids;
    std::vector<int> flipped, changed, moved;
    int oldRoot = -1, oldRootCount = 0;
    double oldKappa = 0.0, oldClock = 0.0, oldLogL = 0.0;
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
        oldLogL = logL;
        oldValid = valid;
        pending = true;
    }
    void undo(int n)
    {
        for (size_t i = undoKids.size(); i-- > 0;) {
            left[undoKids[i].node] = undoKids[i].left;
            right[undoKids[i].node] = undoKids[i].right;
        }
        for (size_t i = undoNodes.size(); i-- > 0;) {
            const NodeUndo &u = undoNodes[i];
            parent[u.node] = u.parent;
            height[u.node] = u.height;
            rate[u.node] = u.rate;
        }
        for (int v : flipped)
            bank[v - n] ^= 1;
        root = oldRoot;
        rootCount = oldRootCount;
        kappa = oldKappa;
        clock = oldClock;
        logL = oldLogL;
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
    double *dMats = nullptr; size_t matsCap = 0;
    double *dPart = nullptr; size_t partCap = 0;
    std::vector<Chain> chains;
    // planner scratch
    std::vector<UpdOp> ops;
    std::vector<int> opOff, stack;
    long long lastOps = 0;
};

namespace {

// transition matrices of the listed nodes: [entry][child][category][16]; internal child row-major,
// tip child transposed (child_term)
__global__ void k2_update_mats(int n, int C, long long nEntries, const UpdOp *ops,
                               const double *kappaPerChain, const double *site, double *mats)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nEntries * 2 * C)
        return;
    const long long k = idx / (2 * C);
    const int j = (int)(idx / C) & 1, c = (int)(idx % C);
    const UpdOp op = ops[k];
    const double *catRates = site, *freqs = site + 2 * C;
    const int child = j ? op.c2 : op.c1;
    const double jointBranchRate = catRates[c] * (j ? op.br2 : op.br1);
    const double distance = (j ? op.dh2 : op.dh1) * jointBranchRate;
    double T[16];
    hky_matrix(kappaPerChain[op.flags >> 8], freqs, distance, T);
    double *M = mats + ((size_t)(k * 2 + j) * C + c) * 16;
    const bool tip = child < n;
#pragma unroll
    for (int e = 0; e < 16; e++)
        M[e] = tip ? T[(e & 3) * 4 + (e >> 2)] : T[e];
}

constexpr int UPD_THREADS = 128;

template <int C>
__global__ void __launch_bounds__(UPD_THREADS)
k2_peel_update(int n, long long P, long long pStride, const uint8_t *states,
               const double *weights, const uint8_t *constMask, const double *site, double pInv,
               const UpdOp *ops, const int *opOff, const double *mats, double *store,
               double *partial, int nTiles)
{
    __shared__ double red[3 * (UPD_THREADS / 32)];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, chain = blockIdx.y;
    const long long p = (long long)tile * UPD_THREADS + tid;        // < pStride
    const int k0 = opOff[chain], k1 = opOff[chain + 1];
    double *out = partial + ((size_t)chain * nTiles + tile) * 3;
    if (k0 == k1) {                     // nothing changed: the host keeps the accepted logL
        if (tid == 0)
            out[0] = out[1] = out[2] = 0.0;
        return;
    }
    constexpr int ROWS = 4 * C + 1;
    const size_t nInt = (size_t)(n - 1);
    auto vec = [&](int node, int bank) -> double * {
        return store + ((((size_t)chain * 2 + bank) * nInt + (size_t)(node - n)) * ROWS) * pStride + p;
    };

    double prev[C][1][4], prevScale = 0.0;
    // child 2 of the entry about to run, fetched one entry ahead
    double nx[C][1][4], nScale = 0.0;
    int nTip[1] = {4};
    auto fetch2 = [&](const UpdOp &op) {
        if (op.c2 < n) {
            nTip[0] = __ldg(states + (size_t)op.c2 * pStride + p);
        } else {
            const double *src = vec(op.c2, (op.flags & UF_C2_BANK) ? 1 : 0);
#pragma unroll
            for (int c = 0; c < C; c++)
#pragma unroll
                for (int i = 0; i < 4; i++)
                    nx[c][0][i] = __ldcg(src + (size_t)(c * 4 + i) * pStride);
            nScale = __ldcg(src + (size_t)(4 * C) * pStride);
        }
    };
    UpdOp op = ops[k0];
    fetch2(op);
#pragma unroll 1
    for (int k = k0; k < k1; k++) {
        const double *M1 = mats + (size_t)k * 2 * C * 16, *M2 = M1 + C * 16;
        // child 1: the previous result, a tip, or (first entry of a path) a resident vector
        int t1[1] = {4};
        double ls = 0.0;
        if (op.flags & UF_PREV) {
            ls = prevScale;
        } else if (op.c1 < n) {
            t1[0] = __ldg(states + (size_t)op.c1 * pStride + p);
        } else {
            const double *src = vec(op.c1, (op.flags & UF_C1_BANK) ? 1 : 0);
#pragma unroll
            for (int c = 0; c < C; c++)
#pragma unroll
                for (int i = 0; i < 4; i++)
                    prev[c][0][i] = __ldcg(src + (size_t)(c * 4 + i) * pStride);
            ls = __ldcg(src + (size_t)(4 * C) * pStride);
        }
        // child 2 (already here); start on the next entry's
        double x2[C][1][4];
        int t2[1] = {nTip[0]};
#pragma unroll
        for (int c = 0; c < C; c++)
#pragma unroll
            for (int i = 0; i < 4; i++)
                x2[c][0][i] = nx[c][0][i];
        if (op.c2 >= n)
            ls += nScale;
        const UpdOp cur = op;
        if (k + 1 < k1) {
            op = ops[k + 1];
            fetch2(op);
        }
        int mxhi = 0;
#pragma unroll
        for (int c = 0; c < C; c++) {
            double a[1][4], bb[1][4];
            child_term<1, 2>(cur.c1 < n, M1 + c * 16, prev[c], t1, a);
            child_term<1, 2>(cur.c2 < n, M2 + c * 16, x2[c], t2, bb);
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
        double *dst = vec(cur.v, (cur.flags & UF_DST_BANK) ? 1 : 0);
#pragma unroll
        for (int c = 0; c < C; c++)
#pragma unroll
            for (int i = 0; i < 4; i++)
                dst[(size_t)(c * 4 + i) * pStride] = prev[c][0][i];
        dst[(size_t)(4 * C) * pStride] = ls;
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
    }
}

template <int C>
int launch_update(mtt_peel_chains *s, const UpdOp *dOps, const int *dOpOff, int nTiles)
{
    mtt_peel *h = s->h;
    mtt_ctx *ctx = h->ctx;
    dim3 grid((unsigned)nTiles, (unsigned)s->K);
    k2_peel_update<C><<<grid, UPD_THREADS, 0, ctx->stream>>>(
        h->nTips, h->nPatterns, h->pStride, h->dStates, h->dWeights, h->dConst, h->dSite, h->pInv,
        dOps, dOpOff, s->dMats, s->dStore, s->dPart, nTiles);
    MTT_CHECK_LAUNCH(ctx);
    return MTT_OK;
}

// Apply the proposal (parent, height, rate) of one chain to its state and append, in post-order,
// the nodes that must be recomputed: an internal node is listed when the branch of a child
// changed (the child's height, clock rate or parent, or the node's own height) or a child is
// listed itself -- the union of the paths from the changes to the root.  `dirty` (optional, N
// flags) names the nodes that MAY differ from the accepted tree; without it every node is
// compared.  Returns false for a malformed tree (the state is rolled back).
bool plan_chain(mtt_peel_chains *s, int k, const int32_t *parent, const double *height,
                const double *rate, double clock, double kappa, const uint8_t *dirty)
{
    const int n = s->h->nTips, N = 2 * n - 1;
    Chain &c = s->chains[k];
    const bool all = !c.valid || kappa != c.kappa || (!rate && clock != c.clock);
    c.begin();
    c.kappa = kappa;
    c.clock = clock;
    std::vector<int> &changed = c.changed;
    changed.clear();
    auto differs = [&](int i) {
        return parent[i] != c.parent[i] || height[i] != c.height[i] ||
               (rate ? rate[i] : clock) != c.rate[i];
    };
    if (all) {
        for (int i = 0; i < N; i++)
            changed.push_back(i);
    } else if (dirty) {
        for (int i = 0; i < N; i++)
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
        c.undo(n);
        return false;
    };
    auto journalKids = [&](int v) { c.undoKids.push_back({v, c.left[v], c.right[v]}); };
    // detach the nodes whose parent changes, then attach them to their new parents
    std::vector<int> &touched = s->stack;
    touched.clear();
    for (int i : changed) {
        const int pp = parent[i];
        if (pp >= N || pp == i || (pp >= 0 && pp < n))
            return fail();
        c.undoNodes.push_back({i, c.parent[i], c.height[i], c.rate[i]});
        if (pp != c.parent[i]) {
            const int old = c.parent[i];
            if (old >= n && old < N) {
                journalKids(old);
                if (c.left[old] == i)
                    c.left[old] = -1;
                else if (c.right[old] == i)
                    c.right[old] = -1;
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
        if (i >= n && (all || height[i] != c.height[i]))
            moved.push_back(i);
        if (pp != c.parent[i]) {
            if (pp >= 0) {
                journalKids(pp);
                if (c.left[pp] < 0)
                    c.left[pp] = i;
                else if (c.right[pp] < 0)
                    c.right[pp] = i;
                else
                    return fail();                  // a third child
                touched.push_back(pp);
            } else {
                c.rootCount++;
                c.root = i;
            }
            c.parent[i] = pp;
        }
        c.height[i] = height[i];
        c.rate[i] = rate ? rate[i] : clock;
    }
    if (c.rootCount != 1 || c.root < n)
        return fail();
    for (int v : touched)
        if (c.left[v] < 0 || c.right[v] < 0)
            return fail();                          // an internal node left with one child
    // walks from the changes up to the root (or to a node already marked in this call); meeting
    // a node marked by the SAME walk means a cycle
    if (c.epoch == 0x7fffffff) {
        c.mark.assign(N, 0);
        c.epoch = 0;
    }
    c.epoch++;
    c.walkId = 0;
    auto walkUp = [&](int u) {
        const int id = ++c.walkId;
        int steps = 0;
        while (u >= 0) {
            if (c.mark[u] == c.epoch) {
                if (c.walk[u] == id)
                    return false;
                return true;
            }
            c.mark[u] = c.epoch;
            c.walk[u] = id;
            u = c.parent[u];
            if (++steps > N)
                return false;
        }
        return true;
    };
    for (int i : changed)
        if (c.parent[i] >= 0 && !walkUp(c.parent[i]))
            return fail();
    for (int i : moved)
        if (!walkUp(i))
            return fail();
    if (c.mark[c.root] != c.epoch)
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
        const int l = c.left[v], r = c.right[v];
        const bool lu = l >= n && c.mark[l] == c.epoch, ru = r >= n && c.mark[r] == c.epoch;
        UpdOp op;
        memset(&op, 0, sizeof(op));
        op.v = v;
        const int c1 = ru ? r : l, c2 = ru ? l : r;
        op.c1 = c1;
        op.c2 = c2;
        op.flags = (k << 8) | ((lu || ru) ? UF_PREV : 0);
        c.bank[v - n] ^= 1;                          // the bank the accepted state does not use
        c.flipped.push_back(v);
        if (c.bank[v - n])
            op.flags |= UF_DST_BANK;
        op.dh1 = c.height[v] - c.height[c1];
        op.dh2 = c.height[v] - c.height[c2];
        op.br1 = c.rate[c1];
        op.br2 = c.rate[c2];
        s->ops.push_back(op);
        if (ru && lu)
            stack.push_back(c2);
        if (lu || ru)
            stack.push_back(c1);
    }
    std::reverse(s->ops.begin() + first, s->ops.end());
    // banks of resident children, known once every listed node has its new bank
    for (size_t e = first; e < s->ops.size(); e++) {
        UpdOp &op = s->ops[e];
        if (op.c1 >= n && c.bank[op.c1 - n])
            op.flags |= UF_C1_BANK;
        if (op.c2 >= n && c.bank[op.c2 - n])
            op.flags |= UF_C2_BANK;
    }
    return true;
}

}  // namespace

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
    if (e != cudaSuccess) {
        delete s;
        return mtt_fail(ctx, MTT_ERR_CUDA, "mtt_peel_chains_create: cudaMalloc(%zu) failed: %s",
                        bytes, cudaGetErrorString(e));
    }
    const int N = 2 * h->nTips - 1;
    s->chains.resize(nChains);
    for (Chain &c : s->chains) {
        c.parent.assign(N, -2);             // -2: no tree yet
        c.left.assign(N, -1);
        c.right.assign(N, -1);
        c.mark.assign(N, 0);
        c.walk.assign(N, 0);
        c.height.assign(N, 0.0);
        c.rate.assign(N, 0.0);
        c.bank.assign(h->nTips - 1, 0);
    }
    *out = s;
    return MTT_OK;
}

int mtt_peel_chains_destroy(mtt_peel_chains *s)
{
    if (!s)
        return MTT_OK;
    cudaSetDevice(s->h->ctx->device);
    cudaStreamSynchronize(s->h->ctx->stream);
    for (void *p : {(void *)s->dStore, (void *)s->dMats, (void *)s->dPart})
        if (p)
            cudaFree(p);
    delete s;
    return MTT_OK;
}

int mtt_peel_chains_update(mtt_peel_chains *s, const int32_t *parent, const double *height,
                           const double *branchRates, const double *kappa,
                           const double *clockRate, const uint8_t *dirty, double *logL)
{
    if (!s)
        return MTT_ERR_INVALID_ARG;
    mtt_peel *h = s->h;
    mtt_ctx *ctx = h->ctx;
    int rc = k2_check_ready(h, kappa == nullptr);
    if (rc != MTT_OK)
        return rc;
    if (!parent || !height || !logL)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "mtt_peel_chains_update: bad arguments");
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    const int n = h->nTips, N = 2 * n - 1, C = h->nCat, K = s->K;
    for (Chain &c : s->chains)
        if (c.pending)
            c.undo(n);                      // an unanswered proposal counts as rejected
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
        const double *rate = branchRates ? branchRates + (size_t)k * N : nullptr;
        kap[k] = kk;
        const size_t before = s->ops.size();
        if (!plan_chain(s, k, parent + (size_t)k * N, height + (size_t)k * N, rate, clock, kk,
                        dirty ? dirty + (size_t)k * N : nullptr)) {
            s->ops.resize(before);
            bad[k] = 1;
        }
        s->opOff[k + 1] = (int)s->ops.size();
    }
    const long long nEntries = (long long)s->ops.size();
    s->lastOps = nEntries;
    if (nEntries > 0) {
        const int nTiles = (int)(h->pStride / UPD_THREADS);
        if ((rc = grow(ctx, s->dMats, s->matsCap, (size_t)nEntries * 2 * C * 16 * 8)) != MTT_OK)
            return rc;
        if ((rc = grow(ctx, s->dPart, s->partCap, (size_t)K * nTiles * 24)) != MTT_OK)
            return rc;
        Xfer x(ctx);
        UpdOp *dOps = nullptr;
        int *dOpOff = nullptr;
        double *dKappa = nullptr, *dLogL = nullptr;
        std::vector<double> got(K);
        x.in(s->ops.data(), s->ops.size(), BUF_C_OPS, &dOps);
        x.in(s->opOff.data(), s->opOff.size(), BUF_C_OPOFF, &dOpOff);
        x.in(kap.data(), kap.size(), BUF_C_KAPPA, &dKappa);
        x.out(got.data(), got.size(), BUF_C_LOGL, &dLogL);
        if ((rc = x.prepare()) != MTT_OK || (rc = x.upload()) != MTT_OK)
            return rc;
        const long long nMat = nEntries * 2 * C;
        k2_update_mats<<<(unsigned)((nMat + 127) / 128), 128, 0, ctx->stream>>>(
            n, C, nEntries, dOps, dKappa, h->dSite, s->dMats);
        MTT_CHECK_LAUNCH(ctx);
        switch (C) {
        case 1: rc = launch_update<1>(s, dOps, dOpOff, nTiles); break;
        case 2: rc = launch_update<2>(s, dOps, dOpOff, nTiles); break;
        case 3: rc = launch_update<3>(s, dOps, dOpOff, nTiles); break;
        case 4: rc = launch_update<4>(s, dOps, dOpOff, nTiles); break;
        case 5: rc = launch_update<5>(s, dOps, dOpOff, nTiles); break;
        case 6: rc = launch_update<6>(s, dOps, dOpOff, nTiles); break;
        case 7: rc = launch_update<7>(s, dOps, dOpOff, nTiles); break;
        case 8: rc = launch_update<8>(s, dOps, dOpOff, nTiles); break;
        default: rc = mtt_fail(ctx, MTT_ERR_TOO_LARGE, "peeling: nCat=%d outside [1,8]", C);
        }
        if (rc != MTT_OK)
            return rc;
        if ((rc = k2_run_finalize(h, K, nTiles, s->dPart, dLogL, nullptr, nullptr, 0)) != MTT_OK)
            return rc;
        if ((rc = x.download()) != MTT_OK)
            return rc;
        for (int k = 0; k < K; k++)
            if (s->opOff[k + 1] > s->opOff[k])
                s->chains[k].logL = got[k];
    }
    for (int k = 0; k < K; k++) {
        Chain &c = s->chains[k];
        if (bad[k]) {
            logL[k] = NAN;                  // malformed tree: NaN, loudly; the chain keeps its state
            continue;
        }
        c.valid = true;                     // (of the proposal; undo() restores the old flag)
        logL[k] = c.logL;                   // unchanged when nothing had to be recomputed
    }
    return MTT_OK;
}

int mtt_peel_chains_accept(mtt_peel_chains *s, const uint8_t *accept)
{
    if (!s)
        return MTT_ERR_INVALID_ARG;
    const int n = s->h->nTips;
    for (int k = 0; k < s->K; k++) {
        Chain &c = s->chains[k];
        if (!c.pending)
            continue;
        if (!accept || accept[k])
            c.commit();
        else
            c.undo(n);
    }
    return MTT_OK;
}

int64_t mtt_peel_chains_last_entries(const mtt_peel_chains *s)
{
    return s ? s->lastOps : 0;
}

}  // extern "C"
