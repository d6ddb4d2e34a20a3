// k5_simulate.cu -- SURVEY 8f rank 4: batched structured-coalescent tree simulation on the device.
//
// Follows the reference's Gillespie initialiser
// evolution/tree/StructuredCoalescentMultiTypeTree.java (simulateTree :191-289,
// updatePropensities :300-327, getNextEvent :352-385, updateTree :395-442) in the formulation of
// the host generator csrc/host/simulator.cpp -- same event loop, same per-tree xoshiro256**
// stream (seed + tree index), same draw order -- so a tree simulated here has the topology,
// types and change lists of the host generator's tree with the same seed, and times equal to
// rounding (the device's log() is not bit-identical to the host's).  Output is the flattened
// tree SoA of include/mtt_b200.h in the owner layout, ready for mtt_sc_logp_owner_batch_dev.
//
//   k5_simulate      one THREAD per tree (the event loop is sequential); per-deme lists of
//                    active lineages in a global scratch.  Run twice: the first pass only
//                    counts the type changes (their number is not known in advance), the second
//                    writes nodes and the raw change list, in simulation order, at chgBase[t].
//   k5_group_changes one warp per tree: stable grouping of the raw changes by owner node
//                    (rank by counting), which is the order the density kernels read.
#include "mtt_common.cuh"

#include <math.h>

#include <algorithm>
#include <numeric>
#include <vector>

namespace {

struct Xoshiro {
    unsigned long long s[4];
    __device__ static unsigned long long splitmix(unsigned long long &x)
    {
        unsigned long long z = (x += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    __device__ explicit Xoshiro(unsigned long long seed)
    {
        for (int i = 0; i < 4; i++)
            s[i] = splitmix(seed);
    }
    __device__ static unsigned long long rotl(unsigned long long x, int k)
    {
        return (x << k) | (x >> (64 - k));
    }
    __device__ unsigned long long next()
    {
        const unsigned long long r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return r;
    }
    __device__ double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    __device__ int below(int n) { return (int)(uniform() * n); }
    __device__ double exponential(double rate) { return -log(1.0 - uniform()) / rate; }
};

struct SimParams {
    long long B;
    int n, d;
    const double *m, *N, *mu;            // device: d*d, d, d
    const uint8_t *leafTypes;            // device [n]
    const double *leafTimes;             // device [n]
    const int *activation;               // device [n]: leaves in order of increasing time (stable)
    unsigned long long seed;
    int *active;                         // scratch [B][d][n]
    int *nChanges;                       // [B]   (count pass)
    // store pass
    double *height;
    int *parent;
    uint8_t *nodeType;
    const long long *chgBase;
    double *rawTime;
    uint8_t *rawType;
    uint16_t *rawNode;
};

template <bool STORE>
__global__ void k5_simulate(const SimParams p)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.B)
        return;
    const int n = p.n, d = p.d, nNodes = 2 * n - 1;
    Xoshiro rng(p.seed + (unsigned long long)t);
    int *act = p.active + (size_t)t * d * n;          // deme i: act[i * n .. i * n + size[i])
    int size[MTT_MAX_DEMES];
    for (int i = 0; i < d; i++)
        size[i] = 0;
    double *H = STORE ? p.height + (size_t)t * nNodes : nullptr;
    int *P = STORE ? p.parent + (size_t)t * nNodes : nullptr;
    uint8_t *T = STORE ? p.nodeType + (size_t)t * nNodes : nullptr;
    long long c = STORE ? p.chgBase[t] : 0;
    if (STORE) {
        for (int i = 0; i < nNodes; i++)
            P[i] = -1;
        for (int i = 0; i < n; i++) {
            H[i] = p.leafTimes[i];
            T[i] = p.leafTypes[i];
        }
    }
    int nextInactive = 0, nextNr = n, nActive = 0, changes = 0;
    double time = 0.0;
    while (nActive > 1 || nextInactive < n) {
        // updatePropensities (:300-327)
        double total = 0.0;
        for (int i = 0; i < d; i++) {
            const double k = (double)size[i];
            total += k * (k - 1) / (2.0 * p.N[i]) + k * p.mu[i];
        }
        // getNextEvent (:352-385)
        double tEvent = INFINITY;
        if (total > 0.0)
            tEvent = time + rng.exponential(total);
        else if (nextInactive >= n)
            break;      // disconnected demes and nothing left to activate: the forest stays
        double U = (total > 0.0) ? rng.uniform() * total : 0.0;
        // activation of the next sampled lineage (:258-276)
        if (nextInactive < n && p.leafTimes[p.activation[nextInactive]] < tEvent) {
            const int leaf = p.activation[nextInactive++];
            time = p.leafTimes[leaf];
            const int ty = p.leafTypes[leaf];
            act[ty * n + size[ty]++] = leaf;
            nActive++;
            continue;
        }
        // updateTree (:395-442): coalescence in i, or migration i -> j, in the reference's order
        bool done = false;
        for (int i = 0; i < d && !done; i++) {
            const double k = (double)size[i];
            const double pc = k * (k - 1) / (2.0 * p.N[i]);
            int *L = act + i * n;
            if (U < pc) {
                const int a = rng.below(size[i]);
                int b = rng.below(size[i] - 1);
                if (b >= a)
                    b++;
                const int daughter = L[a], son = L[b], par = nextNr++;
                if (STORE) {
                    H[par] = tEvent;
                    T[par] = (uint8_t)i;
                    P[daughter] = par;
                    P[son] = par;
                }
                L[a] = par;
                for (int q = b; q + 1 < size[i]; q++)       // erase, order kept
                    L[q] = L[q + 1];
                size[i]--;
                nActive--;
                done = true;
                break;
            }
            U -= pc;
            for (int j = 0; j < d; j++) {
                if (j == i)
                    continue;
                const double pm = k * p.m[i * d + j];
                if (U < pm) {
                    const int a = rng.below(size[i]);
                    const int node = L[a];
                    if (STORE) {
                        p.rawNode[c] = (uint16_t)node;
                        p.rawTime[c] = tEvent;
                        p.rawType[c] = (uint8_t)j;
                        c++;
                    }
                    changes++;
                    for (int q = a; q + 1 < size[i]; q++)
                        L[q] = L[q + 1];
                    size[i]--;
                    act[j * n + size[j]++] = node;
                    done = true;
                    break;
                }
                U -= pm;
            }
        }
        if (!done)
            continue;   // rounding fell through the propensity table: redraw
        time = tEvent;
    }
    if (!STORE)
        p.nChanges[t] = changes;
}

// stable grouping of the raw (simulation-order) changes of every tree by owner node
__global__ void k5_group_changes(long long B, const long long *chgBase, const double *rawTime,
                                 const uint8_t *rawType, const uint16_t *rawNode, double *chgTime,
                                 uint8_t *chgType, uint16_t *chgNode)
{
    const long long t = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= B)
        return;
    const long long c0 = chgBase[t];
    const int M = (int)(chgBase[t + 1] - c0);
    for (int c = lane; c < M; c += 32) {
        const int node = rawNode[c0 + c];
        int rank = 0;
        for (int o = 0; o < M; o++) {
            const int no = rawNode[c0 + o];
            rank += (no < node) || (no == node && o < c);
        }
        chgTime[c0 + rank] = rawTime[c0 + c];
        chgType[c0 + rank] = rawType[c0 + c];
        chgNode[c0 + rank] = (uint16_t)node;
    }
}

}  // namespace

// ---------------------------------------------------------------------------- C-ABI
struct mtt_sim {
    mtt_ctx *ctx = nullptr;
    long long B = 0;
    int n = 0, d = 0;
    long long total = 0;
    double *height = nullptr, *chgTime = nullptr;
    int *parent = nullptr;
    uint8_t *nodeType = nullptr, *chgType = nullptr;
    uint16_t *chgNode = nullptr;
    long long *nodeOff = nullptr, *chgBase = nullptr;
};

extern "C" {

int mtt_sim_trees_create(mtt_ctx *ctx, int64_t B, int nLeaves, int d, const double *m,
                         const double *N, const uint8_t *leafTypes, const double *leafTimes,
                         uint64_t seed, mtt_sim **out)
{
    if (!ctx || !out)
        return MTT_ERR_INVALID_ARG;
    *out = nullptr;
    if (B < 0 || nLeaves < 2 || !m || !N || !leafTypes || !leafTimes)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "mtt_sim_trees_create: bad arguments");
    if (d < 1 || d > MTT_MAX_DEMES)
        return mtt_fail(ctx, MTT_ERR_TOO_LARGE, "mtt_sim_trees_create: d=%d outside [1,%d]", d,
                        MTT_MAX_DEMES);
    if (2 * nLeaves - 1 > 65535)
        return mtt_fail(ctx, MTT_ERR_TOO_LARGE, "mtt_sim_trees_create: more than 32768 tips");
    for (int i = 0; i < nLeaves; i++)
        if (leafTypes[i] >= d)
            return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "mtt_sim_trees_create: leaf type >= d");
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int n = nLeaves, nNodes = 2 * n - 1;

    // model and leaf tables on the device
    std::vector<double> mu(d, 0.0);
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++)
            if (i != j)
                mu[i] += m[i * d + j];
    std::vector<int> activation(n);
    std::iota(activation.begin(), activation.end(), 0);
    std::stable_sort(activation.begin(), activation.end(),
                     [&](int a, int b) { return leafTimes[a] < leafTimes[b]; });
    double *dM, *dN, *dMu, *dLT;
    uint8_t *dLTy;
    int *dAct;
    MTT_CUDA(ctx, cudaMalloc((void **)&dM, (size_t)d * d * 8 + 2 * d * 8 + (size_t)n * 8 + 256));
    dN = dM + d * d;
    dMu = dN + d;
    dLT = dMu + d;
    MTT_CUDA(ctx, cudaMalloc((void **)&dLTy, (size_t)n + 256));
    MTT_CUDA(ctx, cudaMalloc((void **)&dAct, (size_t)n * 4 + 256));
    MTT_CUDA(ctx, cudaMemcpyAsync(dM, m, (size_t)d * d * 8, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dN, N, (size_t)d * 8, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dMu, mu.data(), (size_t)d * 8, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dLT, leafTimes, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dLTy, leafTypes, (size_t)n, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dAct, activation.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));

    mtt_sim *h = new mtt_sim();
    h->ctx = ctx;
    h->B = B;
    h->n = n;
    h->d = d;
    auto fail = [&](int rc) {
        cudaFree(dM);
        cudaFree(dLTy);
        cudaFree(dAct);
        mtt_sim_trees_destroy(h);
        return rc;
    };
    int *dCount = nullptr, *dActive = nullptr;
    // the active-lineage scratch is bounded to ~1 GB: larger batches run in chunks
    const size_t perTree = (size_t)d * n * 4;
    const long long chunk = std::max<long long>(1, std::min<long long>(std::max<long long>(B, 1),
                                                                       (1ll << 30) / (long long)perTree));
    cudaError_t e = cudaMalloc((void **)&dActive, (size_t)chunk * perTree + 256);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&dCount, (size_t)(B + 1) * 4 + 256);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&h->height, (size_t)B * nNodes * 8 + 256);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&h->parent, (size_t)B * nNodes * 4 + 256);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&h->nodeType, (size_t)B * nNodes + 256);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&h->nodeOff, (size_t)(B + 1) * 8 + 256);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&h->chgBase, (size_t)(B + 1) * 8 + 256);
    if (e != cudaSuccess) {
        cudaFree(dActive);
        cudaFree(dCount);
        mtt_fail(ctx, MTT_ERR_CUDA, "mtt_sim_trees_create: cudaMalloc failed: %s", cudaGetErrorString(e));
        return fail(MTT_ERR_CUDA);
    }
    SimParams p{};
    p.n = n; p.d = d; p.m = dM; p.N = dN; p.mu = dMu; p.leafTypes = dLTy; p.leafTimes = dLT;
    p.activation = dAct; p.active = dActive;
    const int TPB = 64;
    // ---- pass 1: number of type changes per tree
    for (long long b0 = 0; b0 < B; b0 += chunk) {
        p.B = std::min(chunk, B - b0);
        p.seed = seed + (unsigned long long)b0;
        p.nChanges = dCount + b0;
        k5_simulate<false><<<(unsigned)((p.B + TPB - 1) / TPB), TPB, 0, s>>>(p);
        ctx->launches++;
    }
    std::vector<int> cnt((size_t)B);
    std::vector<long long> base((size_t)B + 1, 0), noff((size_t)B + 1);
    cudaError_t e2 = cudaMemcpyAsync(cnt.data(), dCount, (size_t)B * 4, cudaMemcpyDeviceToHost, s);
    if (e2 == cudaSuccess)
        e2 = cudaStreamSynchronize(s);
    if (e2 != cudaSuccess) {
        cudaFree(dActive);
        cudaFree(dCount);
        mtt_fail(ctx, MTT_ERR_CUDA, "mtt_sim_trees_create: count pass failed: %s", cudaGetErrorString(e2));
        return fail(MTT_ERR_CUDA);
    }
    for (long long b = 0; b < B; b++) {
        base[b + 1] = base[b] + cnt[b];
        noff[b] = b * nNodes;
    }
    noff[B] = B * nNodes;
    h->total = base[B];
    double *rawTime = nullptr;
    uint8_t *rawType = nullptr;
    uint16_t *rawNode = nullptr;
    const size_t tot = (size_t)h->total;
    e = cudaMalloc((void **)&h->chgTime, tot * 8 + 256);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&h->chgType, tot + 256);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&h->chgNode, tot * 2 + 256);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&rawTime, tot * 8 + 256);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&rawType, tot + 256);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&rawNode, tot * 2 + 256);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(h->chgBase, base.data(), (size_t)(B + 1) * 8, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(h->nodeOff, noff.data(), (size_t)(B + 1) * 8, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        // ---- pass 2: the same streams again, now writing nodes and raw changes
        p.height = h->height; p.parent = h->parent; p.nodeType = h->nodeType;
        p.rawTime = rawTime; p.rawType = rawType; p.rawNode = rawNode;
        for (long long b0 = 0; b0 < B; b0 += chunk) {
            p.B = std::min(chunk, B - b0);
            p.seed = seed + (unsigned long long)b0;
            SimParams q = p;
            q.height += (size_t)b0 * nNodes;
            q.parent += (size_t)b0 * nNodes;
            q.nodeType += (size_t)b0 * nNodes;
            q.chgBase = h->chgBase + b0;
            k5_simulate<true><<<(unsigned)((q.B + TPB - 1) / TPB), TPB, 0, s>>>(q);
            ctx->launches++;
        }
        if (B > 0) {
            k5_group_changes<<<(unsigned)((B + 3) / 4), 128, 0, s>>>(B, h->chgBase, rawTime, rawType,
                                                                     rawNode, h->chgTime, h->chgType,
                                                                     h->chgNode);
            ctx->launches++;
        }
        e = cudaStreamSynchronize(s);
        if (e == cudaSuccess)
            e = cudaGetLastError();
    }
    cudaFree(rawTime);
    cudaFree(rawType);
    cudaFree(rawNode);
    cudaFree(dActive);
    cudaFree(dCount);
    if (e != cudaSuccess) {
        mtt_fail(ctx, MTT_ERR_CUDA, "mtt_sim_trees_create: %s", cudaGetErrorString(e));
        return fail(MTT_ERR_CUDA);
    }
    cudaFree(dM);
    cudaFree(dLTy);
    cudaFree(dAct);
    *out = h;
    return MTT_OK;
}

int mtt_sim_trees_destroy(mtt_sim *h)
{
    if (!h)
        return MTT_OK;
    cudaSetDevice(h->ctx->device);
    void *ptrs[] = {h->height, h->chgTime, h->parent, h->nodeType, h->chgType, h->chgNode,
                    h->nodeOff, h->chgBase};
    for (void *q : ptrs)
        if (q)
            cudaFree(q);
    delete h;
    return MTT_OK;
}

int64_t mtt_sim_trees_total_changes(const mtt_sim *h)
{
    return h ? h->total : 0;
}

int mtt_sim_trees_device(const mtt_sim *h, const int64_t **nodeOff, const double **height,
                         const int32_t **parent, const uint8_t **nodeType,
                         const uint16_t **chgNode, const int64_t **chgBase,
                         const double **chgTime, const uint8_t **chgType)
{
    if (!h)
        return MTT_ERR_INVALID_ARG;
    if (nodeOff) *nodeOff = (const int64_t *)h->nodeOff;
    if (height) *height = h->height;
    if (parent) *parent = h->parent;
    if (nodeType) *nodeType = h->nodeType;
    if (chgNode) *chgNode = h->chgNode;
    if (chgBase) *chgBase = (const int64_t *)h->chgBase;
    if (chgTime) *chgTime = h->chgTime;
    if (chgType) *chgType = h->chgType;
    return MTT_OK;
}

int mtt_sim_trees_export(const mtt_sim *h, int64_t *nodeOff, double *height, int32_t *parent,
                         uint8_t *nodeType, uint16_t *chgNode, int64_t *chgBase, double *chgTime,
                         uint8_t *chgType)
{
    if (!h)
        return MTT_ERR_INVALID_ARG;
    mtt_ctx *ctx = h->ctx;
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t nodes = (size_t)h->B * (2 * h->n - 1), tot = (size_t)h->total;
    if (nodeOff)
        MTT_CUDA(ctx, cudaMemcpyAsync(nodeOff, h->nodeOff, (size_t)(h->B + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (chgBase)
        MTT_CUDA(ctx, cudaMemcpyAsync(chgBase, h->chgBase, (size_t)(h->B + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (height)
        MTT_CUDA(ctx, cudaMemcpyAsync(height, h->height, nodes * 8, cudaMemcpyDeviceToHost, s));
    if (parent)
        MTT_CUDA(ctx, cudaMemcpyAsync(parent, h->parent, nodes * 4, cudaMemcpyDeviceToHost, s));
    if (nodeType)
        MTT_CUDA(ctx, cudaMemcpyAsync(nodeType, h->nodeType, nodes, cudaMemcpyDeviceToHost, s));
    if (tot) {
        if (chgNode)
            MTT_CUDA(ctx, cudaMemcpyAsync(chgNode, h->chgNode, tot * 2, cudaMemcpyDeviceToHost, s));
        if (chgTime)
            MTT_CUDA(ctx, cudaMemcpyAsync(chgTime, h->chgTime, tot * 8, cudaMemcpyDeviceToHost, s));
        if (chgType)
            MTT_CUDA(ctx, cudaMemcpyAsync(chgType, h->chgType, tot, cudaMemcpyDeviceToHost, s));
    }
    MTT_CUDA(ctx, cudaStreamSynchronize(s));
    return MTT_OK;
}

}  // extern "C"
