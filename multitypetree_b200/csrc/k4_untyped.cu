// k4_untyped.cu -- SURVEY 8f rank 1: the untyped-tree (particle) density on sm_100a.
//
// Replaces distributions/StructuredCoalescentUntypedTreeDensity.java:116-207 (calculateLogP),
// :216-290 (colourFirstLineage), :292-316 (colourLineage) and :386-512 (addTypeChanges),
// batched over B untyped trees x nParticles particles.
//
//   k4_particles   one WARP per particle.  The lineages of a particle depend on each other (a
//                  lineage stops at the first ancestor typed by an earlier one), so they run in
//                  sequence inside the warp; the parallelism is particles x trees.  Each
//                  lineage is one call of the uniformization branch sampler shared with
//                  k3_retype (k3_device.cuh: expm, event count, sorted virtual times,
//                  forward-backward types; separately rounded arithmetic, so the sampled typing
//                  is bit-identical to the CPU restatement for the same uniform window).  The
//                  typed tree's events collect in shared memory, are ordered by an exact
//                  (time, insertion index) bitonic sort -- Collections.sort is stable -- and
//                  scanned by lane chunks with per-deme prefix counts.
//   k4_combine     one warp per tree: logP = log(mean_p exp(w_p)) in the reference's
//                  max-shifted form (:193-201).
//
// Latency-bound by design, like k3_retype: the work per particle is a chain of tiny (d <= 16)
// matrix functions.  No tensor cores.
#include "mtt_common.cuh"
#include "k3_device.cuh"

#include <math.h>
#include <string.h>

#include <vector>

namespace {

constexpr int P_WARPS = 4;           // warps (particles) per CTA
constexpr int P_ECAP = 1024;         // most events of one typed tree (nodes + migrations)
constexpr int P_NCAP = 1024;         // most nodes of one tree
// The stage of a warp is sized per call (eCap events, tCap virtual events per lineage, nCap
// nodes) so that small trees leave room for more resident warps; a particle that outgrows it
// reports MTT_ST_SKIPPED / MTT_ST_NEED_MORE_POWERS and the host relaunches with a larger stage.
enum { PK_SAMPLE = 0, PK_COALESCE = 1, PK_MIGRATE = 2 };

struct ParticleParams {
    long long nTrees;
    int nParticles, d, powCap;
    int eCap, tCap, nCap;            // stage capacities (multiples of 8)
    const double *Q, *R, *mu, *pow, *logPow, *logN, *logQ;
    const int *meta;
    const int *model;                // per tree
    const double *N;                 // [nModels][d]
    const long long *nodeOff;
    const double *height;
    const int *parent;
    const uint8_t *leafType;
    const long long *orderOff;       // per tree: start of its nParticles leaf orders
    const int *leafOrder;
    const double *uniforms;
    const long long *uOff;           // per particle
    double *logW;
    int *status;
    long long *uUsed;
    long long partBase, partEnd;     // particles [partBase, partEnd) of this launch
    // typed trees of the particles (k4_particles<true>: mtt_untyped_sample_batch)
    uint8_t *outNodeType;            // particle (t, p): nodeOff[t] * nParticles + p * n
    int chgCap;                      // room per particle in the three change arrays
    int *outNChg;
    double *outChgTime;
    uint8_t *outChgType;
    int *outChgNode;
};

__host__ __device__ __forceinline__ size_t particle_smem_doubles(int d, int eCap, int tCap,
                                                                int nCap)
{
    const int dPad = (d + 1) & ~1;
    return (size_t)8 * d * d + tCap + tCap / 8 + eCap /* evTime */ + eCap / 4 /* ord u16 */ +
           eCap / 4 /* evDesc u16 */ + nCap / 8 /* nodeType */ +
           32 * dPad / 2 /* per-lane counts, int */ + 2 * dPad /* N, muOut */;
}

// TREE: also return the typing itself -- the type of every node and the type changes with the
// node below each (MultiTypeNode.addChange on that node); a lineage without a valid path ends the
// particle with that lineage's status (MultiTypeTreeFromUntypedNewick.java:118-133 starts over).
template <bool TREE>
__global__ void __launch_bounds__(P_WARPS * 32) k4_particles(const ParticleParams p)
{
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long q = p.partBase + (long long)blockIdx.x * P_WARPS + warp;   // particle
    if (q >= p.partEnd)
        return;
    const int d = p.d, dd = d * d;
    __shared__ uint16_t rcAll[P_WARPS][MTT_MAX_DEMES * MTT_MAX_DEMES];
    uint16_t *rc = rcAll[warp];
    fill_rc(rc, d, lane);

    const int eCap = p.eCap, tCap = p.tCap, nCap = p.nCap, dPad = (d + 1) & ~1;
    double *buf = sm + (size_t)warp * particle_smem_doubles(d, eCap, tCap, nCap);
    double *sTimes = buf + 8 * dd;
    uint8_t *sTypes = reinterpret_cast<uint8_t *>(sTimes + tCap);
    double *evTime = sTimes + tCap + tCap / 8;
    uint16_t *ord = reinterpret_cast<uint16_t *>(evTime + eCap);
    uint16_t *evDesc = ord + eCap;
    int8_t *nodeType = reinterpret_cast<int8_t *>(evDesc + eCap);
    int *cnt = reinterpret_cast<int *>(nodeType + nCap) + lane * dPad;
    double *sN = reinterpret_cast<double *>(reinterpret_cast<int *>(nodeType + nCap) + 32 * dPad);
    double *sMuOut = sN + dPad;

    const long long tree = q / p.nParticles;
    const int part = (int)(q - tree * p.nParticles);
    const int mdl = p.model[tree];
    const double *Q = p.Q + (size_t)mdl * dd, *R = p.R + (size_t)mdl * dd;
    const double *lQ = p.logQ + (size_t)mdl * dd;
    const double *tab = p.pow + (size_t)mdl * p.powCap * dd;
    const double *logTab = p.logPow + (size_t)mdl * p.powCap * dd;
    const int size = p.meta[2 * mdl], steady = p.meta[2 * mdl + 1];
    const double mu = p.mu[mdl];
    const long long n0 = p.nodeOff[tree];
    const int n = (int)(p.nodeOff[tree + 1] - n0);
    const int nLeaves = (n + 1) >> 1;
    const double *height = p.height + n0;
    const int *parent = p.parent + n0;
    const int *order = p.leafOrder + p.orderOff[tree] + (long long)part * nLeaves;
    const double *u = p.uniforms + p.uOff[q];
    const long long nU = p.uOff[q + 1] - p.uOff[q];

    for (int i = lane; i < n; i += 32)
        nodeType[i] = (i < nLeaves) ? (int8_t)p.leafType[n0 + i] : (int8_t)-1;   // :124-126
    if (lane < d) {
        sN[lane] = p.N[(size_t)mdl * d + lane];
        double s = 0.0;
        for (int c = 0; c < d; c++)
            if (c != lane)
                s += Q[lane * d + c];
        sMuOut[lane] = s;
    }
    __syncwarp();

    int nEv = 0, st = MTT_ST_OK;
    long long pos = 0;
    double w = 0.0;
    // every lane runs the same scalar control flow; lane 0 owns the shared-memory writes
    auto emit = [&](int kind, int type, int dest, double time) {
        if (lane == 0 && nEv < eCap) {
            evTime[nEv] = time;
            evDesc[nEv] = (uint16_t)(kind << 8 | type << 4 | dest);
        }
        nEv++;
    };
    int nChg = 0;
    auto record = [&](int node, int newType, double time) {      // node.addChange(newType, time)
        if (TREE) {
            if (lane == 0 && nChg < p.chgCap) {
                const long long at = q * p.chgCap + nChg;
                p.outChgTime[at] = time;
                p.outChgType[at] = (uint8_t)newType;
                p.outChgNode[at] = node;
            }
            nChg++;
        }
    };

    for (int li = 0; li < nLeaves && st == MTT_ST_OK; li++) {
        const int leaf = order[li];
        if (li == 0) {
            // ---- colourFirstLineage (:216-290)
            double time = height[leaf];
            int type = nodeType[leaf];
            int next = parent[leaf];
            int thisNode = leaf;
            emit(PK_SAMPLE, type, 0, time);
            double thisLogP = 0.0;
            for (;;) {
                double aTot = 0.0;
                for (int c = 0; c < d; c++)
                    if (c != type)
                        aTot = add_rn(aTot, Q[type * d + c]);
                if (pos >= nU) {
                    st = MTT_ST_NEED_MORE_UNIFORMS;
                    break;
                }
                // Randomizer.nextExponential(aTot) = -1.0 * log(1 - u) / aTot
                const double newTime =
                    add_rn(time, div_ni(mul_rn(-1.0, log_ni(sub_rn(1.0, u[pos++]))), aTot));
                while (next >= 0 && height[next] < newTime) {
                    const double hn = height[next];
                    if (lane == 0)
                        nodeType[next] = (int8_t)type;
                    emit(PK_COALESCE, type, 0, hn);
                    thisLogP = add_rn(thisLogP, mul_rn(-aTot, sub_rn(hn, time)));
                    time = hn;
                    thisNode = next;
                    next = parent[next];
                }
                if (next < 0)
                    break;
                thisLogP = add_rn(thisLogP, mul_rn(-aTot, sub_rn(newTime, time)));
                time = newTime;
                if (pos >= nU) {
                    st = MTT_ST_NEED_MORE_UNIFORMS;
                    break;
                }
                double uu = mul_rn(u[pos++], aTot);                              // :265
                int newType;
                for (newType = 0; newType < d; newType++) {
                    if (newType == type)
                        continue;
                    uu = sub_rn(uu, Q[type * d + newType]);
                    if (uu < 0.0)
                        break;
                }
                if (newType >= d)
                    newType = (type == d - 1) ? d - 2 : d - 1;
                emit(PK_MIGRATE, type, newType, time);
                record(thisNode, newType, time);
                thisLogP = add_rn(thisLogP, lQ[type * d + newType]);
                type = newType;
            }
            w = sub_rn(w, thisLogP);                                             // :134
            __syncwarp();
            continue;
        }
        // ---- colourLineage (:292-316) + addTypeChanges (:386-512)
        const int a = nodeType[leaf];
        emit(PK_SAMPLE, a, 0, height[leaf]);
        int anc = parent[leaf];
        while (nodeType[anc] < 0)
            anc = parent[anc];
        const double t0 = height[leaf], t1 = height[anc];
        const int b = nodeType[anc];
        const BranchSample bs = sample_branch<true>(d, Q, R, tab, logTab, p.logN, size, steady, mu, t0,
                                                    t1, a, b,
                                                    -1, u, nU, pos, buf, rc, sTimes, sTypes,
                                                    nullptr, nullptr, 0, lane, tCap);
        if (bs.st == MTT_ST_NEED_MORE_UNIFORMS || bs.st == MTT_ST_NEED_MORE_POWERS) {
            st = bs.st;
            break;
        }
        if (TREE && bs.st != MTT_ST_OK) {
            st = bs.st;
            break;
        }
        if (bs.st != MTT_ST_OK) {
            // NoValidPath / rubbish / forward-backward failure: the Java returns -Infinity and
            // the weight takes "-= -Infinity" (:136, :311-313, :399-403, :438-441)
            w = sub_rn(w, -INFINITY);
            continue;
        }
        double lp = 0.0;
        int prevType = a, prevNode = leaf, below = leaf;
        double prevTime = t0;
        for (int i = 0; i < bs.nVirt; i++) {                                     // :452-486
            const int ty = bs.types[i];
            if (ty != prevType) {
                const double ti = bs.times[i];
                emit(PK_MIGRATE, prevType, ty, ti);
                while (height[prevNode] < ti) {
                    if (prevNode >= nLeaves) {
                        if (lane == 0)
                            nodeType[prevNode] = (int8_t)prevType;
                        emit(PK_COALESCE, prevType, 0, height[prevNode]);
                    }
                    below = prevNode;
                    prevNode = parent[prevNode];
                }
                record(below, ty, ti);
                lp = add_rn(lp, add_rn(mul_rn(Q[prevType * d + prevType], sub_rn(ti, prevTime)),
                                       lQ[prevType * d + ty]));
                prevType = ty;
                prevTime = ti;
            }
        }
        lp = add_rn(lp, mul_rn(Q[prevType * d + prevType], sub_rn(t1, prevTime)));
        while (height[prevNode] < t1) {                                          // :490-502
            if (prevNode >= nLeaves) {
                if (lane == 0)
                    nodeType[prevNode] = (int8_t)prevType;
                emit(PK_COALESCE, prevType, 0, height[prevNode]);
            }
            prevNode = parent[prevNode];
        }
        lp = sub_rn(lp, isnan(bs.logPba) ? log_ni(bs.Pba) : bs.logPba);
        w = sub_rn(w, lp);                                                       // :136
        __syncwarp();
    }
    if (st == MTT_ST_OK && nEv > eCap)
        st = MTT_ST_SKIPPED;             // more events than the stage holds (host reports it)

    double total = 0.0;
    if (st == MTT_ST_OK) {
        // ---- order the events: exact (time, insertion index), i.e. a stable sort (:143)
        int npad = 32;
        while (npad < nEv)
            npad <<= 1;
        for (int i = lane; i < npad; i += 32)
            ord[i] = (i < nEv) ? (uint16_t)i : (uint16_t)0xffff;
        __syncwarp();
        for (int sz = 2; sz <= npad; sz <<= 1) {
            for (int stride = sz >> 1; stride > 0; stride >>= 1) {
                for (int i = lane; i < (npad >> 1); i += 32) {
                    const int x = 2 * i - (i & (stride - 1));
                    int y = x + stride;
                    if (stride == (sz >> 1))
                        y = (x & ~(sz - 1)) + (sz - 1 - (x & (sz - 1)));
                    const int ix = ord[x], iy = ord[y];
                    const double tx = (ix == 0xffff) ? INFINITY : evTime[ix];
                    const double ty = (iy == 0xffff) ? INFINITY : evTime[iy];
                    if (tx > ty || (tx == ty && ix > iy)) {
                        ord[x] = (uint16_t)iy;
                        ord[y] = (uint16_t)ix;
                    }
                }
                __syncwarp();
            }
        }
        // ---- structured-coalescent density of the typed tree (:145-190): lane chunks
        const int per = (nEv + 31) >> 5;
        const int c0 = min(per * lane, nEv), c1 = min(c0 + per, nEv);
        auto apply = [&](int ds, int *k) {
            const int kind = ds >> 8, type = (ds >> 4) & 15, dest = ds & 15;
            if (kind == PK_COALESCE) {
                k[type] -= 1;
            } else if (kind == PK_MIGRATE) {
                k[type] -= 1;
                k[dest] += 1;
            } else {
                k[type] += 1;
            }
        };
        for (int c = 0; c < d; c++)
            cnt[c] = 0;
        for (int i = c0; i < c1; i++)
            apply(evDesc[ord[i]], cnt);       // event 0 is a SAMPLE: +1 for its type (:145-146)
        for (int c = 0; c < d; c++) {         // exclusive prefix over lanes, per deme
            const int mine = cnt[c];
            int incl = mine;
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL_MASK, incl, o);
                if (lane >= o)
                    incl += v;
            }
            cnt[c] = incl - mine;
        }
        for (int i = c0; i < c1; i++) {
            const int ie = ord[i];
            if (i > 0) {
                const double dt = evTime[ie] - evTime[ord[i - 1]];
                if (dt > 0.0) {                                                  // :154-169
                    double lambda = 0.0;
                    for (int c = 0; c < d; c++) {
                        const int k = cnt[c];
                        lambda += (double)(k * (k - 1)) / (2.0 * sN[c]) + (double)k * sMuOut[c];
                    }
                    total += -dt * lambda;
                }
                const int ds = evDesc[ie];
                const int kind = ds >> 8, type = (ds >> 4) & 15, dest = ds & 15;
                if (kind == PK_COALESCE)
                    total += log(1.0 / sN[type]);
                else if (kind == PK_MIGRATE)
                    total += log(Q[type * d + dest]);
            }
            apply(evDesc[ie], cnt);
        }
        total = warp_sum_down(total);
    }
    if (lane == 0) {
        p.logW[q] = w + total;
        p.status[q] = st;
        p.uUsed[q] = pos;
    }
    if (TREE) {
        __syncwarp();
        uint8_t *o = p.outNodeType + n0 * p.nParticles + (long long)part * n;
        for (int i = lane; i < n; i += 32)
            o[i] = (uint8_t)nodeType[i];           // 255 = not reached (status != MTT_ST_OK)
        if (lane == 0)
            p.outNChg[q] = nChg;
    }
}

// logP = log(sum_p exp(w_p - max) / P) + max  (:193-201); Math.max propagates NaN
__global__ void k4_combine(long long nTrees, int nParticles, const double *logW, double *logP)
{
    const long long tree = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (tree >= nTrees)
        return;
    const double *w = logW + tree * nParticles;
    double mx = -INFINITY;
    bool nan = false;
    for (int i = lane; i < nParticles; i += 32) {
        nan |= isnan(w[i]);
        mx = fmax(mx, w[i]);
    }
    mx = warp_max(mx);
    nan = __any_sync(FULL_MASK, nan);
    if (nan)
        mx = NAN;
    double s = 0.0;
    for (int i = lane; i < nParticles; i += 32)
        s += exp(w[i] - mx);
    s = warp_sum_down(s);
    if (lane == 0)
        logP[tree] = log(s / nParticles) + mx;
}

}  // namespace

namespace {
// Both entries: the density estimate (tr == nullptr) and the typed trees themselves.
struct TreeOut {
    int chgCap;
    uint8_t *nodeType;
    int32_t *nChg;
    double *chgTime;
    uint8_t *chgType;
    int32_t *chgNode;
};
}  // namespace

static int untyped_run(mtt_ctx *ctx, int64_t nTrees, int nParticles, const int32_t *model,
                       const double *N, const int64_t *nodeOff, const double *height,
                       const int32_t *parent, const uint8_t *leafType,
                       const int32_t *leafOrder, const double *uniforms, const int64_t *uOff,
                       double *logP, double *logW, int32_t *status, int64_t *uUsed,
                       const TreeOut *tr)
{
    if (!ctx)
        return MTT_ERR_INVALID_ARG;
    if (ctx->unifModels == 0)
        return mtt_fail(ctx, MTT_ERR_STATE, "mtt_unif_set_models has not been called");
    if (ctx->unifSym)
        return mtt_fail(ctx, MTT_ERR_STATE,
                        "the untyped density uses the unsymmetrised model (sym = 0)");
    if (nTrees < 0 || nParticles < 1 || !N || !nodeOff || !height || !parent || !leafType ||
        !leafOrder || !uniforms || !uOff || (!logP && !tr))
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "mtt_untyped_logp_batch: bad argument");
    if (tr && (tr->chgCap < 0 || !tr->nodeType || !tr->nChg ||
               (tr->chgCap > 0 && (!tr->chgTime || !tr->chgType || !tr->chgNode))))
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "mtt_untyped_sample_batch: bad argument");
    if (nTrees == 0)
        return MTT_OK;
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int d = ctx->unifD, nM = ctx->unifModels;
    const int64_t nPart = nTrees * nParticles, nNodes = nodeOff[nTrees], nU = uOff[nPart];

    // per-tree checks, leaf-order offsets and the longest lineage (bounds the R^k table)
    std::vector<int32_t> hModel(nTrees, 0);
    std::vector<long long> hOrderOff(nTrees + 1, 0);
    double maxMuL = 0.0;
    int64_t maxNodes = 0;
    for (int64_t t = 0; t < nTrees; t++) {
        const int64_t n0 = nodeOff[t], n = nodeOff[t + 1] - n0;
        maxNodes = n > maxNodes ? n : maxNodes;
        if (n < 3 || (n & 1) == 0 || n > P_NCAP)
            return mtt_fail(ctx, MTT_ERR_TOO_LARGE,
                            "untyped density: tree %lld has %lld nodes (supported: odd, 3..%d)",
                            (long long)t, (long long)n, P_NCAP);
        if (model) {
            if (model[t] < 0 || model[t] >= nM)
                return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "tree %lld: model out of range",
                                (long long)t);
            hModel[t] = model[t];
        }
        hOrderOff[t + 1] = hOrderOff[t] + (long long)nParticles * ((n + 1) / 2);
        double lo = height[n0], hi = height[n0];
        for (int64_t i = 0; i < n; i++) {
            lo = height[n0 + i] < lo ? height[n0 + i] : lo;
            hi = height[n0 + i] > hi ? height[n0 + i] : hi;
        }
        const double muL = ctx->unifMu[hModel[t]] * (hi - lo);
        if (muL > maxMuL)
            maxMuL = muL;
    }
    if (!(maxMuL < 1e7))
        return mtt_fail(ctx, MTT_ERR_TOO_LARGE, "mu*L = %g is too large", maxMuL);
    int K = (int)(maxMuL + 10.0 * sqrt(maxMuL) + 32.0);
    if (K > U_TCAP)
        K = U_TCAP;          // a lineage with more virtual events than the stage is reported below

    int *dModel = (int *)mtt_buf(ctx, BUF_P_MODEL, nTrees * 4);
    double *dN = (double *)mtt_buf(ctx, BUF_P_N, (size_t)nM * d * 8);
    long long *dNodeOff = (long long *)mtt_buf(ctx, BUF_P_NODEOFF, (nTrees + 1) * 8);
    double *dHeight = (double *)mtt_buf(ctx, BUF_P_HEIGHT, nNodes * 8);
    int *dParent = (int *)mtt_buf(ctx, BUF_P_PARENT, nNodes * 4);
    uint8_t *dLeafType = (uint8_t *)mtt_buf(ctx, BUF_P_LEAFTYPE, nNodes);
    long long *dOrderOff = (long long *)mtt_buf(ctx, BUF_P_ORDEROFF, (nTrees + 1) * 8);
    int *dOrder = (int *)mtt_buf(ctx, BUF_P_ORDER, hOrderOff[nTrees] * 4);
    double *dU = (double *)mtt_buf(ctx, BUF_P_U, nU * 8);
    long long *dUOff = (long long *)mtt_buf(ctx, BUF_P_UOFF, (nPart + 1) * 8);
    double *dLogW = (double *)mtt_buf(ctx, BUF_P_LOGW, nPart * 8);
    int *dStatus = (int *)mtt_buf(ctx, BUF_P_STATUS, nPart * 4);
    long long *dUsed = (long long *)mtt_buf(ctx, BUF_P_USED, nPart * 8);
    double *dLogP = (double *)mtt_buf(ctx, BUF_P_LOGP, nTrees * 8);
    uint8_t *dOutType = nullptr, *dChgType = nullptr;
    int *dNChg = nullptr, *dChgNode = nullptr;
    double *dChgTime = nullptr;
    const size_t nChgRoom = tr ? (size_t)nPart * tr->chgCap : 0;
    if (tr) {
        dOutType = (uint8_t *)mtt_buf(ctx, BUF_P_OUTTYPE, (size_t)nNodes * nParticles);
        dNChg = (int *)mtt_buf(ctx, BUF_P_NCHG, nPart * 4);
        dChgTime = (double *)mtt_buf(ctx, BUF_P_CHGTIME, nChgRoom * 8 + 8);
        dChgType = (uint8_t *)mtt_buf(ctx, BUF_P_CHGTYPE, nChgRoom + 8);
        dChgNode = (int *)mtt_buf(ctx, BUF_P_CHGNODE, nChgRoom * 4 + 8);
        if (!dOutType || !dNChg || !dChgTime || !dChgType || !dChgNode)
            return MTT_ERR_CUDA;
    }
    if (!dModel || !dN || !dNodeOff || !dHeight || !dParent || !dLeafType || !dOrderOff ||
        !dOrder || !dU || !dUOff || !dLogW || !dStatus || !dUsed || !dLogP)
        return MTT_ERR_CUDA;
    MTT_CUDA(ctx, cudaMemcpyAsync(dModel, hModel.data(), nTrees * 4, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dN, N, (size_t)nM * d * 8, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dNodeOff, nodeOff, (nTrees + 1) * 8, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dHeight, height, nNodes * 8, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dParent, parent, nNodes * 4, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dLeafType, leafType, nNodes, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dOrderOff, hOrderOff.data(), (nTrees + 1) * 8,
                                  cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dOrder, leafOrder, hOrderOff[nTrees] * 4,
                                  cudaMemcpyHostToDevice, s));
    // The uniforms are nearly all of the input.  Large calls send them in chunks of particles on
    // a second stream, each chunk's particles starting as soon as its uniforms have landed
    // (first attempt only: a retry with larger stages finds everything on the device).
    const int nChunks = (nPart >= 4096) ? 8 : 1;
    if (nChunks == 1)
        MTT_CUDA(ctx, cudaMemcpyAsync(dU, uniforms, nU * 8, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dUOff, uOff, (nPart + 1) * 8, cudaMemcpyHostToDevice, s));
    if (nChunks > 1) {
        if (!ctx->copyStream)
            MTT_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
        while ((int)ctx->events.size() < nChunks + 1) {
            cudaEvent_t ev;
            MTT_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            ctx->events.push_back(ev);
        }
    }

    // stage capacities: start from the tree size / the expected number of virtual events and
    // grow on demand
    int nCap = 8 * (int)((maxNodes + 7) / 8);
    int eCap = 64;
    while (eCap < 2 * maxNodes && eCap < P_ECAP)
        eCap <<= 1;
    int tCap = 32;
    while (tCap < K && tCap < U_TCAP)
        tCap <<= 1;
    const unsigned blocks = (unsigned)((nPart + P_WARPS - 1) / P_WARPS);
    std::vector<int32_t> hSt(nPart);
    for (int attempt = 0;; attempt++) {
        int rc = mtt_unif_ensure_powers(ctx, K);
        if (rc != MTT_OK)
            return rc;
        const size_t smem = (size_t)P_WARPS * particle_smem_doubles(d, eCap, tCap, nCap) * 8;
        void (*kernel)(const ParticleParams) = tr ? k4_particles<true> : k4_particles<false>;
        MTT_CUDA(ctx, mtt_fn_smem(ctx, kernel, smem));
        ParticleParams p;
        p.nTrees = nTrees; p.nParticles = nParticles; p.d = d; p.powCap = ctx->unifPowCap;
        p.eCap = eCap; p.tCap = tCap; p.nCap = nCap;
        p.Q = (const double *)ctx->bufs[BUF_U_Q].p; p.R = (const double *)ctx->bufs[BUF_U_R].p;
        p.logQ = (const double *)ctx->bufs[BUF_U_LOGQ].p;
        p.mu = (const double *)ctx->bufs[BUF_U_MU].p;
        p.pow = (const double *)ctx->bufs[BUF_U_POW].p;
        p.logPow = (const double *)ctx->bufs[BUF_U_LOGPOW].p;
        p.logN = (const double *)ctx->bufs[BUF_U_LOGN].p;
        p.meta = (const int *)ctx->bufs[BUF_U_META].p;
        p.model = dModel; p.N = dN; p.nodeOff = dNodeOff; p.height = dHeight; p.parent = dParent;
        p.leafType = dLeafType; p.orderOff = dOrderOff; p.leafOrder = dOrder; p.uniforms = dU;
        p.uOff = dUOff; p.logW = dLogW; p.status = dStatus; p.uUsed = dUsed;
        p.outNodeType = dOutType; p.chgCap = tr ? tr->chgCap : 0; p.outNChg = dNChg;
        p.outChgTime = dChgTime; p.outChgType = dChgType; p.outChgNode = dChgNode;
        if (attempt == 0 && nChunks > 1) {
            // (the buffers of the previous call are free: every call ends with a synchronise)
            MTT_CUDA(ctx, cudaEventRecord(ctx->events[nChunks], s));
            MTT_CUDA(ctx, cudaStreamWaitEvent(ctx->copyStream, ctx->events[nChunks], 0));
            const int64_t per = (nPart + nChunks - 1) / nChunks;
            for (int c = 0; c < nChunks; c++) {
                const int64_t q0 = std::min<int64_t>(nPart, per * c);
                const int64_t q1 = std::min<int64_t>(nPart, per * (c + 1));
                if (q1 <= q0)
                    continue;
                const int64_t u0 = uOff[q0], u1 = uOff[q1];
                if (u1 > u0)
                    MTT_CUDA(ctx, cudaMemcpyAsync(dU + u0, uniforms + u0, (size_t)(u1 - u0) * 8,
                                                  cudaMemcpyHostToDevice, ctx->copyStream));
                MTT_CUDA(ctx, cudaEventRecord(ctx->events[c], ctx->copyStream));
                MTT_CUDA(ctx, cudaStreamWaitEvent(s, ctx->events[c], 0));
                p.partBase = q0;
                p.partEnd = q1;
                kernel<<<(unsigned)((q1 - q0 + P_WARPS - 1) / P_WARPS), P_WARPS * 32, smem, s>>>(p);
                MTT_CHECK_LAUNCH(ctx);
            }
        } else {
            p.partBase = 0;
            p.partEnd = nPart;
            kernel<<<blocks, P_WARPS * 32, smem, s>>>(p);
            MTT_CHECK_LAUNCH(ctx);
        }
        MTT_CUDA(ctx, cudaMemcpyAsync(hSt.data(), dStatus, nPart * 4, cudaMemcpyDeviceToHost, s));
        MTT_CUDA(ctx, cudaStreamSynchronize(s));
        bool morePow = false, moreEv = false;
        for (int64_t i = 0; i < nPart; i++) {
            morePow |= hSt[i] == MTT_ST_NEED_MORE_POWERS;
            moreEv |= hSt[i] == MTT_ST_SKIPPED;
        }
        if (!morePow && !moreEv)
            break;
        if ((moreEv && eCap >= P_ECAP) || attempt >= 12)
            return mtt_fail(ctx, MTT_ERR_TOO_LARGE,
                            "untyped density: a particle needs more than %d events", P_ECAP);
        if (morePow && K >= U_TCAP && tCap >= U_TCAP)
            return mtt_fail(ctx, MTT_ERR_TOO_LARGE,
                            "untyped density: a lineage needs more than %d virtual events",
                            U_TCAP);
        if (moreEv)
            eCap <<= 1;
        if (morePow) {              // the R^k table or the virtual-event stage was too short
            K = (2 * K + 64 > U_TCAP) ? U_TCAP : 2 * K + 64;
            while (tCap < K && tCap < U_TCAP)
                tCap <<= 1;
        }
    }
    if (logP) {
        k4_combine<<<(unsigned)((nTrees + 3) / 4), 128, 0, s>>>(nTrees, nParticles, dLogW, dLogP);
        MTT_CHECK_LAUNCH(ctx);
        MTT_CUDA(ctx, cudaMemcpyAsync(logP, dLogP, nTrees * 8, cudaMemcpyDeviceToHost, s));
    }
    if (tr) {
        MTT_CUDA(ctx, cudaMemcpyAsync(tr->nodeType, dOutType, (size_t)nNodes * nParticles,
                                      cudaMemcpyDeviceToHost, s));
        MTT_CUDA(ctx, cudaMemcpyAsync(tr->nChg, dNChg, nPart * 4, cudaMemcpyDeviceToHost, s));
        if (nChgRoom) {
            MTT_CUDA(ctx, cudaMemcpyAsync(tr->chgTime, dChgTime, nChgRoom * 8,
                                          cudaMemcpyDeviceToHost, s));
            MTT_CUDA(ctx, cudaMemcpyAsync(tr->chgType, dChgType, nChgRoom,
                                          cudaMemcpyDeviceToHost, s));
            MTT_CUDA(ctx, cudaMemcpyAsync(tr->chgNode, dChgNode, nChgRoom * 4,
                                          cudaMemcpyDeviceToHost, s));
        }
    }
    if (logW)
        MTT_CUDA(ctx, cudaMemcpyAsync(logW, dLogW, nPart * 8, cudaMemcpyDeviceToHost, s));
    if (uUsed)
        MTT_CUDA(ctx, cudaMemcpyAsync(uUsed, dUsed, nPart * 8, cudaMemcpyDeviceToHost, s));
    MTT_CUDA(ctx, cudaStreamSynchronize(s));
    if (status)
        memcpy(status, hSt.data(), nPart * 4);
    return MTT_OK;
}

extern "C" int mtt_untyped_logp_batch(mtt_ctx *ctx, int64_t nTrees, int nParticles,
                                      const int32_t *model, const double *N,
                                      const int64_t *nodeOff, const double *height,
                                      const int32_t *parent, const uint8_t *leafType,
                                      const int32_t *leafOrder, const double *uniforms,
                                      const int64_t *uOff, double *logP, double *logW,
                                      int32_t *status, int64_t *uUsed)
{
    if (ctx && !logP)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "mtt_untyped_logp_batch: logP is NULL");
    return untyped_run(ctx, nTrees, nParticles, model, N, nodeOff, height, parent, leafType,
                       leafOrder, uniforms, uOff, logP, logW, status, uUsed, nullptr);
}

extern "C" int mtt_untyped_sample_batch(mtt_ctx *ctx, int64_t nTrees, int nParticles,
                                        const int32_t *model, const double *N,
                                        const int64_t *nodeOff, const double *height,
                                        const int32_t *parent, const uint8_t *leafType,
                                        const int32_t *leafOrder, const double *uniforms,
                                        const int64_t *uOff, int chgCap, uint8_t *nodeType,
                                        int32_t *nChg, double *chgTime, uint8_t *chgType,
                                        int32_t *chgNode, double *logW, int32_t *status,
                                        int64_t *uUsed)
{
    const TreeOut tr = {chgCap, nodeType, nChg, chgTime, chgType, chgNode};
    return untyped_run(ctx, nTrees, nParticles, model, N, nodeOff, height, parent, leafType,
                       leafOrder, uniforms, uOff, nullptr, logW, status, uUsed, &tr);
}
