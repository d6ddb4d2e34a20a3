// k1_density.cu -- piece 1: batched structured-coalescent tree density on sm_100a.
//
// Replaces distributions/StructuredCoalescentTreeDensity.java:86-147 (calculateLogP),
// :153-248 (updateEventSequence) and evolution/tree/MultiTypeTree.java:396-429 (isValid).
//
// The reference rebuilds a time-ordered event list with an O(E x active lineages) scan and
// then walks it.  Here the flattened tree (height / nodeType / chgCount / chgTime / chgType,
// include/mtt_b200.h) is turned into events, ordered and scanned inside one kernel:
//
//   k1_density_warp   one WARP per tree, up to 1022 events (stages of 256 / 384 / 512 / 1024
//                     slots).  Persistent warps, 24 resident per SM up to 384 slots; lane 0
//                     stages the tree's five array slices into the warp's shared-memory stage
//                     with 1-D bulk async copies (TMA engine, mbarrier completion).  Events are
//                     addressed by stage SLOT; slots without an event are harmless twins of the
//                     last node, so no pass carries a bound.  Trees up to 254 events (config 2)
//                     stay in registers: times read once, a bitonic network over 256 packed
//                     32-bit keys (23-bit quantised time | 9-bit slot; shuffles across lanes),
//                     a warp prefix sum of packed (biased) 16-bit lineage counts, a fully
//                     unrolled scan of each lane's 8 consecutive intervals that also verifies
//                     the exact fp64 order (quantisation ties are repaired, rarely).  Larger
//                     trees: sorted runs of 256 in shared memory, merge path, rolled scan.
//   k1_density_cta    one CTA per tree beyond that (configs 3 and 4, up to 16384 events), fed
//                     from the queue the warp kernel fills: register-sorted runs per warp,
//                     merge-path levels across the CTA, an equalised time grid for the keys,
//                     the same scan after a CTA-wide prefix sum.
//
// fp64 throughout; no atomics on the result.  HBM traffic per tree is the input arrays once
// (<= 12 B per event) + 8 B out.  The staging alone runs at the HBM roofline; at the 2-deme /
// 100-tip size the kernel as a whole is bound by the ordering network (ALU pipe / issue slots):
// see DESIGN.md section 4.1 and profiles/.
#include "mtt_common.cuh"

#include <math.h>
#include <stdlib.h>

namespace {

constexpr int KIND_SAMPLE = 0, KIND_COALESCE = 1, KIND_MIGRATE = 2, KIND_NONE = 3;

struct K1Params {
    long long B;
    int d;
    long long nModels;
    const double *m, *N;
    const long long *nodeOff;
    const double *height;
    const int *parent;
    const uint8_t *nodeType;
    const uint16_t *chgNode;     // tree-local number of the node that owns each change
    const long long *chgBase;
    const double *chgTime;
    const uint8_t *chgType;
    int checkValidity;
    double *logP;
    uint8_t *valid;
    int *queue;        // [0] = count, [1..] = tree indices for the block kernel
};

// ---- per-model tables (shared memory) ------------------------------------------------
// two-deme path: everything an event does, looked up by its class = 2 * kind + from
// (SAMPLE 0/1, COALESCE 2/3, MIGRATE 4/5 with to = 1 - from, NONE 6/7).  An event's
// descriptor is the byte offset of its record, 16 * class: records of different classes lie in
// different shared-memory banks, so a warp's look-ups never conflict.
struct ClsRec {
    double term;                     // log(1/N) / log(m) event term
    uint32_t dk0hi, dk1hi;           // high words of the doubles -1.0 / 0.0 / +1.0: change of
                                     // the lineage count of deme 0 / 1 (their low words are 0)
};
constexpr int D8_SHIFT = 4;          // log2(sizeof(ClsRec))
constexpr int D8_MIGRATE = 4 << D8_SHIFT, D8_NONE = 6 << D8_SHIFT;

struct ModelTab {
    double inv2N[MTT_MAX_DEMES];     // 1 / (2 N_c)
    double mu[MTT_MAX_DEMES];        // sum_{c' != c} m(c,c')
    double logInvN[MTT_MAX_DEMES];   // log(1.0 / N_c)      (StructuredCoalescentTreeDensity.java:130)
    double logM[MTT_MAX_DEMES * MTT_MAX_DEMES];   // log(m(c,c'))   (:136)
    alignas(16) ClsRec cls[8];
    uint32_t clsDelta[8];            // the count changes as a packed pair of 16-bit fields
};

// cooperative fill by `nthr` threads with index `tid`
__device__ __forceinline__ void fill_model_tab(ModelTab *t, const double *m, const double *N,
                                               int d, int tid, int nthr)
{
    for (int c = tid; c < d; c += nthr) {
        double Nc = N[c];
        t->inv2N[c] = 1.0 / (2.0 * Nc);
        t->logInvN[c] = log(1.0 / Nc);
        double s = 0.0;
        for (int cp = 0; cp < d; cp++)
            if (cp != c)
                s += m[c * d + cp];
        t->mu[c] = s;
    }
    for (int e = tid; e < d * d; e += nthr)
        t->logM[e] = (e / d == e % d) ? log(1.0 / N[e / d]) : log(m[e]);   // diagonal: coalescence
    if (d <= 2) {
        // a single deme still reads the second deme's coefficients (as zeros)
        for (int c = d + tid; c < 2; c += nthr) {
            t->inv2N[c] = 0.0;
            t->mu[c] = 0.0;
        }
        for (int c = tid; c < 8; c += nthr) {
            const int kind = c >> 1, from = c & 1, to = 1 - from;
            double term = 0.0, dFrom = 0.0, dTo = 0.0;       // count changes of `from` / `to`
            if (from < d) {
                if (kind == KIND_SAMPLE) {
                    dFrom = 1.0;
                } else if (kind == KIND_COALESCE) {
                    dFrom = -1.0;
                    term = log(1.0 / N[from]);
                } else if (kind == KIND_MIGRATE && to < d) {
                    dFrom = -1.0;
                    dTo = 1.0;
                    term = log(m[from * d + to]);
                }
            }
            const double dk0 = from == 0 ? dFrom : dTo, dk1 = from == 0 ? dTo : dFrom;
            t->cls[c].term = term;
            t->cls[c].dk0hi = (uint32_t)__double2hiint(dk0);
            t->cls[c].dk1hi = (uint32_t)__double2hiint(dk1);
            t->clsDelta[c] = (uint32_t)((int)dk0 + (int)dk1 * 65536);
        }
    }
}

// lineage counts packed as 16-bit fields, 4 demes per 64-bit word
template <int NW>
struct Counts {
    unsigned long long w[NW];
};

template <int NW>
__device__ __forceinline__ void cnt_add(Counts<NW> &c, int deme, long long sgn)
{
    unsigned long long inc = (unsigned long long)sgn << (16 * (deme & 3));
    if (NW == 1) {
        c.w[0] += inc;
    } else {
#pragma unroll
        for (int i = 0; i < NW; i++)
            if ((deme >> 2) == i)
                c.w[i] += inc;
    }
}

// NW == 0: the two-deme specialisation of the warp path -- one 32-bit word, two 16-bit fields
template <>
struct Counts<0> {
    uint32_t w32;
};
template <int NW>
__device__ __forceinline__ void cnt_zero(Counts<NW> &c)
{
#pragma unroll
    for (int i = 0; i < NW; i++)
        c.w[i] = 0;
}
template <>
__device__ __forceinline__ void cnt_zero<0>(Counts<0> &c)
{
    c.w32 = 0;
}

// branch-free variant for converged code: SAMPLE enters `from`, COALESCE leaves `from`,
// MIGRATE leaves `from` and enters `to`, NONE does nothing
template <int NW>
__device__ __forceinline__ void cnt_apply_bf(Counts<NW> &c, uint32_t desc)
{
    const int kind = (desc >> 10) & 3, from = (desc >> 5) & 31, to = desc & 31;
    const bool leave = (kind == KIND_COALESCE) | (kind == KIND_MIGRATE);
    const bool enter = (kind == KIND_SAMPLE) | (kind == KIND_MIGRATE);
    cnt_add<NW>(c, from, -(long long)leave);
    cnt_add<NW>(c, (kind == KIND_MIGRATE) ? to : from, (long long)enter);
}

template <int NW>
__device__ __forceinline__ void cnt_apply(Counts<NW> &c, uint32_t desc)
{
    int kind = (desc >> 10) & 3, from = (desc >> 5) & 31, to = desc & 31;
    if (kind == KIND_SAMPLE)
        cnt_add<NW>(c, from, 1);
    else if (kind == KIND_COALESCE)
        cnt_add<NW>(c, from, -1);
    else if (kind == KIND_MIGRATE) {
        cnt_add<NW>(c, from, -1);
        cnt_add<NW>(c, to, 1);
    }
}

// lambda = sum_c k_c (k_c - 1) / (2 N_c) + sum_c k_c sum_{c'!=c} m(c,c')     (:109-123)
template <int NW>
__device__ __forceinline__ double cnt_lambda(const Counts<NW> &c, const ModelTab *t, int d)
{
    double lambda = 0.0;
#pragma unroll
    for (int i = 0; i < NW; i++) {
#pragma unroll
        for (int f = 0; f < 4; f++) {
            int deme = i * 4 + f;
            if (deme < d) {
                int k = (int)((c.w[i] >> (16 * f)) & 0xffffu);
                double kd = (double)k;
                lambda += (kd * (kd - 1.0)) * t->inv2N[deme] + kd * t->mu[deme];
            }
        }
    }
    return lambda;
}

__device__ __forceinline__ double event_term(uint32_t desc, const ModelTab *t, int d)
{
    int kind = (desc >> 10) & 3, from = (desc >> 5) & 31, to = desc & 31;
    if (kind == KIND_COALESCE)
        return t->logInvN[from];
    if (kind == KIND_MIGRATE)      // m(i,i) = 0 in the reference (SCMigrationModel.java:253-255)
        return from == to ? -INFINITY : t->logM[from * d + to];
    return 0.0;
}

// =====================================================================================
// warp-per-tree kernel
// =====================================================================================
// A staged tree is addressed by SLOT: the index of an event's time in the stage's array of
// doubles.  Node e sits in slot e, change c in slot cs + c (cs = n, n + 1 or n + 2: the bulk
// copies land on 16-byte boundaries); `desc[slot]` describes the event.  Every slot of the
// padded range [0, Epad) that holds no event -- the gap before the changes, the tail, ignored
// root-branch changes -- is turned into a twin of the last node: the same time, a descriptor that
// does nothing.  A twin adds exactly 0 to the sum (its interval has length 0), so nothing
// downstream needs a sentinel, a bound or a predicate: the sort takes Epad keys, the scan
// takes Epad events.  The number of twins depends only on n + M, never on the alignment of
// the tree inside the batch, so a tree's logP does not depend on where it sits.
//
// Trees with n + M + 2 <= 256 (config 2) take the register path: times are read once (range
// and sort keys come from the same registers), the ordered keys never leave the registers
// (each lane owns the 8 consecutive positions the network leaves it with) and the interval
// scan is fully unrolled.  Larger trees go through shared memory: sorted runs of 256 keys,
// merge path, rolled scan loops.
//
// The kernel body must stay resident in the 32 KB instruction cache: an earlier fully
// unrolled version (13k SASS instructions) spent most of its cycles in instruction-fetch
// stalls.
// warps per CTA: 8, or 5 in the 1024-event stage so that two CTAs (10 warps) fit an SM
__host__ __device__ constexpr int warps_per_cta(int EMAX) { return EMAX > 512 ? 5 : 8; }
// sort key = quantised time | slot: 9 slot bits up to 512 slots, 10 in the 1024-event stage
// (one bit less of time resolution; ties are repaired exactly)
__host__ __device__ constexpr int idx_bits(int EMAX) { return EMAX > 512 ? 10 : 9; }
// a tree fits a stage when n + M + 2 <= EMAX (up to two alignment slots before the changes)
constexpr int SLOT_SLACK = 2;

template <int NW>
struct DescOf {
    using T = uint16_t;      // kind | from | to
};
template <>
struct DescOf<0> {
    using T = uint8_t;       // 32 * class
};

template <bool HAS_PARENT, int EMAX>
struct WarpStage {
    alignas(16) unsigned char time[8 * EMAX + 64];   // heights | change times
    alignas(16) unsigned char type[EMAX + 64];       // node types | change types
    alignas(16) unsigned char cnt[2 * EMAX + 32];    // change owners; later the order list
    alignas(16) unsigned char par[HAS_PARENT ? 4 * EMAX + 32 : 16];
};

template <int NW, bool HAS_PARENT, int EMAX>
struct WarpSmem {
    WarpStage<HAS_PARENT, EMAX> st;       // one staged tree; latency is hidden by the other
                                          // resident warps, not by prefetching
    // sorted runs of the multi-run path; with parent[] also the validity pass's byte flags
    alignas(16) uint32_t keys[(EMAX > 256 || HAS_PARENT) ? EMAX : 4];
    alignas(8) unsigned long long bar;
    alignas(16) unsigned char desc[(NW == 0 ? 1 : 2) * (EMAX + 32)];
    alignas(16) uint32_t keys2[EMAX > 512 ? EMAX : 4];   // second buffer of the 4-run merge
};

struct TreeSpan {
    long long n0, c0;
    int n, M;
};

__device__ __forceinline__ TreeSpan load_span(const K1Params &p, long long t)
{
    TreeSpan s;
    if (t < p.B) {
        s.n0 = p.nodeOff[t];
        s.n = (int)(p.nodeOff[t + 1] - s.n0);
        s.c0 = p.chgBase[t];
        s.M = (int)(p.chgBase[t + 1] - s.c0);
    } else {
        s.n0 = s.c0 = 0;
        s.n = s.M = 0;
    }
    return s;
}

// A slice [src, src+bytes) is staged as its 16 B-aligned superset: `a` leading pad bytes,
// `sz` total bytes (a multiple of 16; 0 for an empty slice).
__device__ __forceinline__ void slice_of(const void *src, uint32_t bytes, uint32_t &a,
                                         uint32_t &sz)
{
    a = (uint32_t)((uintptr_t)src & 15);
    sz = bytes ? ((a + bytes + 15u) & ~15u) : 0u;
}

// byte layout of a staged tree inside a WarpStage
struct StageLayout {
    uint32_t hOff, cOff, tOff, ctOff, kOff, pOff;     // where the data starts
    uint32_t szH, szC, szT, szCT, szK, szP;           // bulk-copy sizes
};

template <bool HAS_PARENT>
__device__ __forceinline__ StageLayout make_layout(const K1Params &p, const TreeSpan &s)
{
    StageLayout L;
    uint32_t a;
    slice_of(p.height + s.n0, 8u * s.n, a, L.szH);
    L.hOff = a;
    slice_of(p.chgTime + s.c0, 8u * s.M, a, L.szC);
    L.cOff = L.szH + a;
    slice_of(p.nodeType + s.n0, (uint32_t)s.n, a, L.szT);
    L.tOff = a;
    slice_of(p.chgType + s.c0, (uint32_t)s.M, a, L.szCT);
    L.ctOff = L.szT + a;
    slice_of(p.chgNode + s.c0, 2u * s.M, a, L.szK);
    L.kOff = a;
    L.pOff = 0;
    L.szP = 0;
    if (HAS_PARENT) {
        slice_of(p.parent + s.n0, 4u * s.n, a, L.szP);
        L.pOff = a;
    }
    return L;
}

// one elected lane: arm the stage's mbarrier and launch the bulk copies
template <bool HAS_PARENT, int EMAX>
__device__ __forceinline__ void issue_tree(const K1Params &p, const TreeSpan &s,
                                           const StageLayout &L,
                                           WarpStage<HAS_PARENT, EMAX> *st, uint64_t *bar)
{
#if defined(MTT_K1_EXP_ONECOPY)      // experiment: only the heights are staged (results are garbage)
    mbar_arrive_expect_tx(bar, L.szH);
    if (L.szH)
        bulk_g2s(st->time, (const unsigned char *)(p.height + s.n0) - L.hOff, L.szH, bar);
    return;
#elif defined(MTT_K1_EXP_NOCOPY)     // experiment: nothing is staged
    mbar_arrive_expect_tx(bar, 0);
    return;
#endif
#ifdef MTT_K1_LANE_STAGE      // measured neutral (500 vs 504 M trees/s): off
    // The TMA engine takes ~50 cycles per request whatever its size: only the two time arrays
    // (85 % of the bytes) and parent[] go through it, the three small arrays are fetched by
    // the lanes (stage_small).
    mbar_arrive_expect_tx(bar, L.szH + L.szC + L.szP);
    if (L.szH)
        bulk_g2s(st->time, (const unsigned char *)(p.height + s.n0) - L.hOff, L.szH, bar);
    if (L.szC)
        bulk_g2s(st->time + L.szH, (const unsigned char *)(p.chgTime + s.c0) - (L.cOff - L.szH),
                 L.szC, bar);
    if (HAS_PARENT && L.szP)
        bulk_g2s(st->par, (const unsigned char *)(p.parent + s.n0) - L.pOff, L.szP, bar);
#else
    mbar_arrive_expect_tx(bar, L.szH + L.szC + L.szT + L.szCT + L.szK + L.szP);
    if (L.szH)
        bulk_g2s(st->time, (const unsigned char *)(p.height + s.n0) - L.hOff, L.szH, bar);
    if (L.szC) {
        bulk_g2s(st->time + L.szH, (const unsigned char *)(p.chgTime + s.c0) - (L.cOff - L.szH),
                 L.szC, bar);
        bulk_g2s(st->type + L.szT, (const unsigned char *)(p.chgType + s.c0) - (L.ctOff - L.szT),
                 L.szCT, bar);
        bulk_g2s(st->cnt, (const unsigned char *)(p.chgNode + s.c0) - L.kOff, L.szK, bar);
    }
    if (L.szT)
        bulk_g2s(st->type, (const unsigned char *)(p.nodeType + s.n0) - L.tOff, L.szT, bar);
    if (HAS_PARENT && L.szP)
        bulk_g2s(st->par, (const unsigned char *)(p.parent + s.n0) - L.pOff, L.szP, bar);
#endif
}

// The whole warp: node types, change types and change owners as 16-byte chunks, one per lane
// and round, into the same places the bulk copies would put them.
template <bool HAS_PARENT, int EMAX>
__device__ __forceinline__ void stage_small(const K1Params &p, const TreeSpan &s,
                                            const StageLayout &L,
                                            WarpStage<HAS_PARENT, EMAX> *st, int lane)
{
    const uint32_t cT = L.szT >> 4, cCT = L.szCT >> 4, cK = L.szK >> 4;
    const unsigned char *gT = (const unsigned char *)(p.nodeType + s.n0) - L.tOff;
    const unsigned char *gCT = (const unsigned char *)(p.chgType + s.c0) - (L.ctOff - L.szT);
    const unsigned char *gK = (const unsigned char *)(p.chgNode + s.c0) - L.kOff;
#pragma unroll 1
    for (uint32_t c = lane; c < cT + cCT + cK; c += 32) {
        const unsigned char *src;
        unsigned char *dst;
        if (c < cT) {
            src = gT + 16u * c;
            dst = st->type + 16u * c;
        } else if (c < cT + cCT) {
            src = gCT + 16u * (c - cT);
            dst = st->type + L.szT + 16u * (c - cT);
        } else {
            src = gK + 16u * (c - cT - cCT);
            dst = st->cnt + 16u * (c - cT - cCT);
        }
        *reinterpret_cast<uint4 *>(dst) = __ldg(reinterpret_cast<const uint4 *>(src));
    }
}

// L2 prefetch of a slice's 16 B-aligned superset (what the bulk copy will read)
__device__ __forceinline__ void prefetch_slice(const void *src, uint32_t bytes)
{
    if (bytes) {
        const uint32_t a = (uint32_t)((uintptr_t)src & 15);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((const unsigned char *)src - a),
                     "r"((a + bytes + 15u) & ~15u)
                     : "memory");
    }
}

// one elected lane: pull the next tree of this warp into L2 while the current one is worked on
template <bool HAS_PARENT>
__device__ __forceinline__ void prefetch_tree(const K1Params &p, const TreeSpan &s)
{
    prefetch_slice(p.height + s.n0, 8u * s.n);
    prefetch_slice(p.nodeType + s.n0, (uint32_t)s.n);
    prefetch_slice(p.chgTime + s.c0, 8u * s.M);
    prefetch_slice(p.chgType + s.c0, (uint32_t)s.M);
    prefetch_slice(p.chgNode + s.c0, 2u * s.M);
    if (HAS_PARENT)
        prefetch_slice(p.parent + s.n0, 4u * s.n);
}

__device__ __forceinline__ void ce_u32(uint32_t &a, uint32_t &b)
{
    const uint32_t lo = min(a, b), hi = max(a, b);
    a = lo;
    b = hi;
}

// Bitonic sort of 256 keys held 8 per lane; position = 8 * lane + r.  Every compare-exchange
// is "lower position keeps the minimum": each merge starts with the reflected partner,
// followed by ordinary half cleaners (no direction flags).
__device__ __forceinline__ void warp_sort256(uint32_t (&v)[8], int lane)
{
    constexpr int EPL = 8;
#pragma unroll
    for (int size = 2; size <= EPL; size <<= 1) {
#pragma unroll
        for (int r = 0; r < EPL; r++) {
            const int q = r ^ (size - 1);
            if (q > r)
                ce_u32(v[r], v[q]);
        }
#pragma unroll
        for (int s = size >> 2; s >= 1; s >>= 1) {
#pragma unroll
            for (int r = 0; r < EPL; r++)
                if ((r & s) == 0)
                    ce_u32(v[r], v[r | s]);
        }
    }
#pragma unroll
    for (int lanes = 2; lanes <= 32; lanes <<= 1) {
        {   // reflected partner: lane ^ (lanes - 1), register EPL-1-r
            const bool lower = (lane & (lanes >> 1)) == 0;
            uint32_t t[EPL];
#pragma unroll
            for (int r = 0; r < EPL; r++)
                t[r] = __shfl_xor_sync(FULL_MASK, v[EPL - 1 - r], lanes - 1);
#pragma unroll
            for (int r = 0; r < EPL; r++)
                v[r] = lower ? min(v[r], t[r]) : max(v[r], t[r]);
        }
#pragma unroll
        for (int s = lanes >> 2; s >= 1; s >>= 1) {      // half cleaners across lanes
            const bool lower = (lane & s) == 0;
#pragma unroll
            for (int r = 0; r < EPL; r++) {
                const uint32_t t = __shfl_xor_sync(FULL_MASK, v[r], s);
                v[r] = lower ? min(v[r], t) : max(v[r], t);
            }
        }
#pragma unroll
        for (int s = EPL >> 1; s >= 1; s >>= 1) {        // half cleaners inside the lane
#pragma unroll
            for (int r = 0; r < EPL; r++)
                if ((r & s) == 0)
                    ce_u32(v[r], v[r | s]);
        }
    }
}

// The warp path keeps its running counts BIASED by CNT_BIAS per 16-bit field.  Events with
// exactly equal times are processed in slot order, which need not be a valid lineage order
// (a coalescence may be seen before the same-time migration that brings its child into the
// deme), so a count can be -1 transiently; the bias keeps such a field from borrowing from
// its neighbour, and lambda -- a polynomial in the counts -- is updated consistently through
// the transient.  An interval only contributes at dt > 0, when the counts are true again.
constexpr int CNT_BIAS = 0x4000;
constexpr unsigned long long CNT_BIAS_WORD = 0x4000400040004000ull;

template <int NW>
__device__ __forceinline__ int cnt_get(const Counts<NW> &c, int deme)
{
    unsigned long long w = c.w[0];
    if (NW > 1) {
#pragma unroll
        for (int i = 1; i < NW; i++)
            if ((deme >> 2) == i)
                w = c.w[i];
    }
    return (int)((w >> (16 * (deme & 3))) & 0xffffu) - CNT_BIAS;
}

// count deltas of one ordered event added to the lane's packed sum
template <int NW>
__device__ __forceinline__ void cnt_accumulate(Counts<NW> &c, uint32_t ds, const ModelTab *tab)
{
    if constexpr (NW == 0)
        c.w32 += *reinterpret_cast<const uint32_t *>(
            reinterpret_cast<const unsigned char *>(tab->clsDelta) + (ds >> 2));
    else
        cnt_apply_bf<NW>(c, ds);
}

// Quantisation of event times for the sort keys.  The range comes from the high words of the
// times (min / max as integers: three reductions cover every sign combination); lo is rounded
// down and hi up by less than 2^-20 relative.  q(t) = low word of (t - lo) * scale + 2^52.
struct Quant {
    double lo, scale;
};

struct HiWordRange {
    int smin, smax;
    uint32_t umax;
    __device__ __forceinline__ void init()
    {
        smin = 0x7fffffff;
        smax = (int)0x80000000;
        umax = 0u;
    }
    __device__ __forceinline__ void add(double t)
    {
        const int hw = __double2hiint(t);
        smin = min(smin, hw);
        smax = max(smax, hw);
        umax = max(umax, (uint32_t)hw);
    }
};

template <int IDX_BITS>
__device__ __forceinline__ Quant make_quant(HiWordRange r)
{
    constexpr double QMAX = (double)((1u << (32 - IDX_BITS)) - 2u);
    r.smin = __reduce_min_sync(FULL_MASK, r.smin);
    r.smax = __reduce_max_sync(FULL_MASK, r.smax);
    r.umax = __reduce_max_sync(FULL_MASK, r.umax);
    // a negative time exists <=> umax has the sign bit: the most negative one has the largest
    // unsigned high word; otherwise the smallest signed one is the minimum.  The maximum is
    // the largest signed word unless every time is negative (then the smallest word).
    const int hwLo = ((int)r.umax < 0) ? (int)r.umax : r.smin;
    const int hwHi = (r.smax >= 0) ? r.smax : r.smin;
    Quant q;
    q.lo = __hiloint2double(hwLo, hwLo < 0 ? -1 : 0);
    const double hi = __hiloint2double(hwHi, hwHi < 0 ? 0 : -1);
    const double range = hi - q.lo;
    q.scale = 0.0;
    if (range > 1e-30 && range < 1e30)
        // reciprocal in single precision, biased so that (hi - lo) * scale <= QMAX
        q.scale = (double)__frcp_rd(__double2float_ru(range)) * (QMAX * (1.0 - 1.0 / 262144.0));
    else if (range > 1e-300 && range < INFINITY)
        q.scale = QMAX * (1.0 - 1.0 / 262144.0) / range;
    return q;
}

template <int IDX_BITS>
__device__ __forceinline__ uint32_t quant_key(const Quant &q, double t, uint32_t slot)
{
    const double v = fma(t - q.lo, q.scale, 4503599627370496.0);     // + 2^52
    return ((uint32_t)__double2loint(v) << IDX_BITS) | slot;
}

// One staged tree (n + M + 2 <= EMAX).  Returns the warp-reduced logP in lane 0; `okOut` is
// the validity verdict, `badOut` "malformed input" (all lanes).  The stage's chgNode region,
// dead after phase A, receives the time-ordered slot list when one is needed.
template <int NW, bool HAS_PARENT, int EMAX>
__device__ __forceinline__ double process_tree(WarpSmem<NW, HAS_PARENT, EMAX> *ws,
                                            const StageLayout &L, int n, int M,
                                            const ModelTab *tab, int d, int lane, bool &okOut,
                                            bool &badOut)
{
    using DT = typename DescOf<NW>::T;
    constexpr int IDX_BITS = idx_bits(EMAX);
    constexpr uint32_t IDX_MASK = (1u << IDX_BITS) - 1u;
    WarpStage<HAS_PARENT, EMAX> *st = &ws->st;
    const int E = n + M;
    const int nLeaves = (n + 1) >> 1;
    double *tm = reinterpret_cast<double *>(st->time + L.hOff);          // indexed by slot
    const int cs = (int)((L.cOff - L.hOff) >> 3);                        // slot of change 0
    const int Stot = cs + M;
    const uint8_t *sT = st->type + L.tOff;
    const uint8_t *sCT = st->type + L.ctOff;
    const uint16_t *sO = reinterpret_cast<const uint16_t *>(st->cnt + L.kOff);   // change owners
    const int *sP = reinterpret_cast<const int *>(st->par + L.pOff);
    // descriptors by slot; the two-deme bytes are shifted like the staged node types so that
    // four of them are written per 32-bit word of types
    DT *desc = reinterpret_cast<DT *>(ws->desc) + (NW == 0 ? L.tOff : 0u);
    uint16_t *order = reinterpret_cast<uint16_t *>(st->cnt);             // written after phase A
    const unsigned char *clsBytes = reinterpret_cast<const unsigned char *>(tab->cls);
    const unsigned char *dltBytes = reinterpret_cast<const unsigned char *>(tab->clsDelta);

    constexpr int E2 = EMAX < 512 ? EMAX : 512;      // stage size of the two-run case
    const bool fast = E + SLOT_SLACK <= 256;
    int Epad = fast ? 256 : ((EMAX > 512 && E + SLOT_SLACK > 512) ? EMAX : E2);

#ifdef MTT_K1_EXP_NOPHASEA
    if (fast) {
        uint32_t key[8];
        double t[8];
        HiWordRange hr;
        hr.init();
#pragma unroll
        for (int k = 0; k < 8; k++) {
            t[k] = tm[lane + 32 * k];
            hr.add(t[k]);
        }
        const Quant qz = make_quant<IDX_BITS>(hr);
#pragma unroll
        for (int k = 0; k < 8; k++)
            key[k] = quant_key<IDX_BITS>(qz, t[k], (uint32_t)(lane + 32 * k));
        okOut = true; badOut = false;
        return (double)(key[0] + key[3] + key[7] + key[1] + key[2] + key[4] + key[5] + key[6]);
    }
#endif
#ifdef MTT_K1_EXP_NOKEYS
    okOut = true; badOut = false;
    return tm[lane];
#endif
    // ---- phase A: event descriptors.  Every change names the node that owns it (chgNode),
    //      changes of one node are adjacent and ordered child -> parent: the first change of a
    //      branch leaves the node's type, any other one the type its predecessor entered.
    bool bad = false, ok = true, selfMig = false;
    if constexpr (NW == 0) {
        const uint32_t *tyW = reinterpret_cast<const uint32_t *>(st->type);
        uint32_t *dW = reinterpret_cast<uint32_t *>(ws->desc);
        const int nb = (int)L.tOff + nLeaves;                // first internal node's byte
        const int wb = nb >> 2;
        const uint32_t part = 0x20202020u << (8 * (nb & 3));
        const int nW = ((int)L.tOff + n + 3) >> 2;
        uint32_t seen = 0;
#pragma unroll 1
        for (int w = lane; w < nW; w += 32) {
            const uint32_t t4 = tyW[w];
            seen |= t4;
            const uint32_t add = (w > wb) ? 0x20202020u : ((w == wb) ? part : 0u);
            dW[w] = ((t4 & 0x01010101u) << D8_SHIFT) + add;         // SAMPLE / COALESCE of deme t
        }
        // a type >= 2: possibly a neighbour's byte in the first or last word -- look again
        if (__any_sync(FULL_MASK, (seen & (d >= 2 ? 0xfefefefeu : 0xffffffffu)) != 0u)) {
#pragma unroll 1
            for (int e = lane; e < n; e += 32)
                bad |= (int)sT[e] >= d;
        }
    } else {
#pragma unroll 2
        for (int e = lane; e < n; e += 32) {
            const int type = sT[e];
            bad |= type >= d;
            desc[e] = (DT)((e < nLeaves) ? (KIND_SAMPLE << 10 | type << 5)
                                         : (KIND_COALESCE << 10 | type << 5 | type));
        }
    }
    __syncwarp();
    constexpr DT DESC_NONE = (NW == 0) ? (DT)D8_NONE : (DT)(KIND_NONE << 10);
    if (!HAS_PARENT) {
#pragma unroll 1
        for (int c = lane; c < M; c += 32) {
            const int node = sO[c];
            const bool first = (c == 0) || ((int)sO[c - 1] != node);
            const bool oob = node >= n;                      // owner outside the tree
            const int from = first ? (int)sT[oob ? 0 : node] : (int)sCT[c - 1];
            const int to = sCT[c];
            bad |= (from >= d) | (to >= d);
            ok = ok && !oob;
            // a change into the type it leaves: the reference's rate m(i,i) is 0, the term
            // log(0) (SCMigrationModel.java:253-255); the counts do not move
            selfMig |= (from == to) && !oob;
            desc[cs + c] = (from == to || oob) ? DESC_NONE   // (an owner-less change is ignored)
                           : (NW == 0)  ? (DT)(D8_MIGRATE + ((from & 1) << D8_SHIFT))
                                        : (DT)(KIND_MIGRATE << 10 | (from & 31) << 5 | (to & 31));
        }
    } else {
        // with parent[]: the validity chain of MultiTypeTree.java:396-429 rides along
        uint8_t *hasChg = reinterpret_cast<uint8_t *>(ws->keys);     // free until the sort
        for (int i = lane * 16; i < n; i += 512)
            *reinterpret_cast<uint4 *>(hasChg + i) = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
#pragma unroll 1
        for (int c = lane; c < M; c += 32) {
            const int node = sO[c];
            const bool oob = node >= n;
            const int nd = oob ? 0 : node;
            const bool first = (c == 0) || ((int)sO[c - 1] != node);
            const bool last = (c == M - 1) || ((int)sO[c + 1] != node);
            const int par = sP[nd];
            const bool isRoot = par < 0;
            const int from = first ? (int)sT[nd] : (int)sCT[c - 1];
            const int to = sCT[c];
            bad |= (from >= d) | (to >= d) | (par >= n);
            selfMig |= (from == to) && !isRoot && !oob;
            // root-branch changes are ignored (StructuredCoalescentTreeDensity.java:161-162)
            desc[cs + c] = (oob || isRoot || from == to)
                               ? DESC_NONE
                               : ((NW == 0) ? (DT)(D8_MIGRATE + ((from & 1) << D8_SHIFT))
                                            : (DT)(KIND_MIGRATE << 10 | (from & 31) << 5 |
                                                   (to & 31)));
            const double tc = tm[cs + c];
            const double before = first ? tm[nd] : tm[cs + c - 1];
            ok = ok && !oob && !(before > tc);               // :404-405
            if (last && !isRoot && !oob && par < n) {
                ok = ok && !(tc > tm[par]);                  // :401, :408-409
                ok = ok && (to == (int)sT[par]);             // :421-422
            }
            if (first && !oob)
                hasChg[nd] = 1;
        }
        __syncwarp();
#pragma unroll 1
        for (int e = lane; e < n; e += 32) {                 // branches without changes
            const int par = sP[e];
            bad |= par >= n;
            if (par >= 0 && par < n && !hasChg[e]) {
                ok = ok && !(tm[e] > tm[par]);
                ok = ok && (sT[e] == sT[par]);
            }
        }
        __syncwarp();
    }
    // ---- twins in every slot without an event.  They copy the LAST node (the root in BEAST's
    //      usual numbering, else some internal node): its time is shared with no other event,
    //      so practically never does a real event have to be carried past the block of twins
    //      by the exact-order repair (which a crowded leaf time would make likely).
    {
        const double t0 = (n > 0) ? tm[n - 1] : 0.0;
        __syncwarp();
        if (lane < cs - n) {
            tm[n + lane] = t0;
            desc[n + lane] = DESC_NONE;
        }
#pragma unroll 1
        for (int s = Stot + lane; s < Epad; s += 32) {
            tm[s] = t0;
            desc[s] = DESC_NONE;
        }
        if (HAS_PARENT) {
#pragma unroll 1
            for (int c = lane; c < M; c += 32)
                if (desc[cs + c] == DESC_NONE)
                    tm[cs + c] = t0;
        }
    }
    {   // one reduction for the three verdicts
        const uint32_t f = __reduce_or_sync(FULL_MASK, (bad ? 1u : 0u) | (ok ? 0u : 2u) |
                                                           (selfMig ? 4u : 0u));
        bad = f & 1u;
        ok = !(f & 3u);
        selfMig = f & 4u;
    }
    okOut = ok;
    badOut = bad;
    __syncwarp();

    double acc;
    if (fast) {
        // =========================== register path: 256 slots, 8 per lane ================
        uint32_t key[8];
        {
            double t[8];
            HiWordRange hr;
            hr.init();
#pragma unroll
            for (int k = 0; k < 8; k++) {
                t[k] = tm[lane + 32 * k];
                hr.add(t[k]);
            }
            const Quant qz = make_quant<IDX_BITS>(hr);
#pragma unroll
            for (int k = 0; k < 8; k++)
                key[k] = quant_key<IDX_BITS>(qz, t[k], (uint32_t)(lane + 32 * k));
        }
#ifndef MTT_K1_EXP_NOSORT       // (experiment: cost of everything but the ordering network)
        warp_sort256(key, lane);
#endif
#pragma unroll
        for (int r = 0; r < 8; r++)
            key[r] &= IDX_MASK;                              // slots, in time order
#ifdef MTT_K1_EXP_NOSCAN        // (experiment: cost of everything but the scan)
        return (double)(key[0] + key[3] + key[7]);
#endif
        bool repaired = false;
        for (;;) {
            // time of the last event before my chunk = the previous lane's last event
            // (lane 0: my own first event, dt = 0)
            const double xLast = tm[key[7]], xFirst = tm[key[0]];
            double prevT = __shfl_up_sync(FULL_MASK, xLast, 1);
            if (lane == 0)
                prevT = xFirst;
            int invAcc = 0;
            acc = 0.0;
            if constexpr (NW == 0) {
                uint32_t dsc[8];
                uint32_t local = 0;
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    dsc[r] = desc[key[r]];
                    local += *reinterpret_cast<const uint32_t *>(dltBytes + (dsc[r] >> 2));
                }
                uint32_t incl = local;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(FULL_MASK, incl, o);
                    if (lane >= o)
                        incl += v;
                }
                const uint32_t b32 = incl - local + 0x40004000u;      // biased, see CNT_BIAS
                double k0 = (double)((int)(b32 & 0xffffu) - CNT_BIAS);
                double k1 = (double)((int)(b32 >> 16) - CNT_BIAS);
                // lambda = sum_c k_c (k_c - 1) / (2 N_c) + k_c mu_c   (:109-123), as
                // k_c (k_c / (2 N_c) + (mu_c - 1 / (2 N_c)))
                const double i0 = tab->inv2N[0], i1 = tab->inv2N[1];
                const double b0 = tab->mu[0] - i0, b1 = tab->mu[1] - i1;
                double lambda = k0 * (k0 * i0 + b0) + k1 * (k1 * i1 + b1);
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    const double x = (r == 0) ? xFirst : (r == 7) ? xLast : tm[key[r]];
                    const uint4 rec = *reinterpret_cast<const uint4 *>(clsBytes + dsc[r]);
                    const double dt = x - prevT;
                    invAcc |= __double2hiint(dt);            // sign bit: dt < 0
                    acc = fma(-dt, lambda, acc);             // dt = 0 adds nothing  (:108)
                    prevT = x;
                    // log(1/N) or log(m)  (:127-137)
                    const double term = __hiloint2double((int)rec.y, (int)rec.x);
                    acc += term;
                    k0 += __hiloint2double((int)rec.z, 0);
                    k1 += __hiloint2double((int)rec.w, 0);
                    lambda = k0 * (k0 * i0 + b0) + k1 * (k1 * i1 + b1);
                    if (r & 1)       // keep the scheduler from hoisting all 32 loads (spills)
                        asm volatile("" ::: "memory");
                }
                if (lane == 0)                               // event 0 is skipped (:101)
                    acc -= reinterpret_cast<const ClsRec *>(clsBytes + dsc[0])->term;
            } else {
                Counts<NW> local;
                cnt_zero<NW>(local);
#pragma unroll
                for (int r = 0; r < 8; r++)
                    cnt_apply_bf<NW>(local, desc[key[r]]);
                Counts<NW> base;
#pragma unroll
                for (int i = 0; i < NW; i++) {
                    unsigned long long incl = local.w[i];
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned long long v = shfl_up_u64(incl, o);
                        if (lane >= o)
                            incl += v;
                    }
                    base.w[i] = incl - local.w[i] + CNT_BIAS_WORD;
                }
                double lambda = 0.0;
#pragma unroll 1
                for (int c = 0; c < d; c++) {
                    const double kd = (double)cnt_get<NW>(base, c);
                    lambda += (kd * (kd - 1.0)) * tab->inv2N[c] + kd * tab->mu[c];
                }
                bool first = (lane == 0);                    // event 0 is skipped (:101)
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    const double x = (r == 0) ? xFirst : (r == 7) ? xLast : tm[key[r]];
                    const uint32_t ds = desc[key[r]];
                    const int kind = (ds >> 10) & 3, from = (ds >> 5) & 31, to = ds & 31;
                    const double dt = x - prevT;
                    invAcc |= __double2hiint(dt);
                    acc = fma(-dt, lambda, acc);
                    prevT = x;
                    // Branch-free (lanes hold different event kinds): a lineage leaves deme
                    // `from` (COALESCE, MIGRATE) and / or enters deme `a` (SAMPLE, MIGRATE).
                    // MIGRATE has a != from, so both updates can use the counts before the
                    // event.  NONE does neither.
                    const bool leave = (kind == KIND_COALESCE) | (kind == KIND_MIGRATE);
                    const bool enter = (kind == KIND_SAMPLE) | (kind == KIND_MIGRATE);
                    const int a = (kind == KIND_MIGRATE) ? to : from;
                    const double term = tab->logM[from * d + to];   // log(1/N) or log(m)
                    acc += (leave && !first) ? term : 0.0;
                    const double kl = (double)cnt_get<NW>(base, from),
                                 ke = (double)cnt_get<NW>(base, a);
                    const double dl = (2.0 * (kl - 1.0)) * tab->inv2N[from] + tab->mu[from];
                    const double de = (2.0 * ke) * tab->inv2N[a] + tab->mu[a];
                    lambda += (enter ? de : 0.0) - (leave ? dl : 0.0);
                    cnt_add<NW>(base, from, -(long long)leave);
                    cnt_add<NW>(base, a, (long long)enter);
                    first = false;
                }
            }
#ifdef MTT_K1_EXP_NOSORT
            break;
#endif
            if (repaired || !__any_sync(FULL_MASK, invAcc < 0))
                break;                                       // (NaN times cannot be ordered)
            repaired = true;
            // rare: quantisation ties out of exact order -> odd-even transposition on the
            // exact times, through the order list in shared memory
#pragma unroll
            for (int r = 0; r < 8; r++)
                order[8 * lane + r] = (uint16_t)key[r];
            __syncwarp();
            bool swapped = true;
            while (swapped) {
                swapped = false;
#pragma unroll 1
                for (int parity = 0; parity < 2; parity++) {
#pragma unroll 1
                    for (int a = parity + 2 * lane; a + 1 < 256; a += 64) {
                        const int ia = order[a], ib = order[a + 1];
                        if (tm[ia] > tm[ib]) {
                            order[a] = (uint16_t)ib;
                            order[a + 1] = (uint16_t)ia;
                            swapped = true;
                        }
                    }
                    __syncwarp();
                }
                swapped = __any_sync(FULL_MASK, swapped);
            }
#pragma unroll
            for (int r = 0; r < 8; r++)
                key[r] = order[8 * lane + r];
        }
        return selfMig ? -INFINITY : warp_sum_down(acc);
    }

    if constexpr (EMAX > 256) {
        // =========================== shared-memory path: sorted runs + merge path =========
        uint32_t *keys = ws->keys, *keys2 = ws->keys2;
        constexpr int LB = E2 - 256;                     // capacity of the second run
        asm volatile("" : "+r"(Epad));       // opaque trip count: keeps the sort loop from being peeled
        Quant qz;
        {
            HiWordRange hr;
            hr.init();
#pragma unroll 2
            for (int s = lane; s < Epad; s += 32)
                hr.add(tm[s]);
            qz = make_quant<IDX_BITS>(hr);
        }
#pragma unroll 1
        for (int base = 0; base < Epad; base += 256) {
            uint32_t key[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int s = base + lane + 32 * k;
                // slots past Epad only exist in the two-run stage (Epad = 384): sort last
                key[k] = (s < Epad) ? quant_key<IDX_BITS>(qz, tm[s], (uint32_t)s) : 0xffffffffu;
            }
            warp_sort256(key, lane);
            if (Epad > 512 || base == 0 || 8 * lane < LB) {
                // two runs: only the first LB keys of the second are real
                uint4 *dst = reinterpret_cast<uint4 *>(keys + base + 8 * lane);
                dst[0] = make_uint4(key[0], key[1], key[2], key[3]);
                dst[1] = make_uint4(key[4], key[5], key[6], key[7]);
            }
        }
        __syncwarp();
        // packed count deltas of the positions this lane will own ([per*lane, per*lane+per)),
        // accumulated while the merged keys pass through registers
        Counts<NW> local;
        cnt_zero<NW>(local);
        {
            // merge path: lane produces outputs [per*lane, per*lane+per) of two ascending runs
            auto merge_keys = [&](const uint32_t *A, const uint32_t *B, uint32_t *out) {
                // two 256-key runs -> 512 keys (first level of the four-run case)
                const int diag = 16 * lane;
                int a0 = max(0, diag - 256), a1 = min(diag, 256);
                while (a0 < a1) {
                    const int mid = (a0 + a1) >> 1;
                    if (A[mid] <= B[diag - 1 - mid])
                        a0 = mid + 1;
                    else
                        a1 = mid;
                }
                int i = a0, j = diag - a0;
                uint32_t ka = (i < 256) ? A[i] : 0xffffffffu, kb = (j < 256) ? B[j] : 0xffffffffu;
#pragma unroll 1
                for (int q = 0; q < 16; q++) {
                    const bool takeA = ka <= kb;
                    out[diag + q] = takeA ? ka : kb;
                    if (takeA) {
                        i++;
                        ka = (i < 256) ? A[i] : 0xffffffffu;
                    } else {
                        j++;
                        kb = (j < 256) ? B[j] : 0xffffffffu;
                    }
                }
            };
            const uint32_t *A = keys, *B = keys + 256;
            int la = 256, lb = Epad - 256;
            if (EMAX > 512 && Epad > 512) {
                merge_keys(keys, keys + 256, keys2);
                merge_keys(keys + 512, keys + 768, keys2 + 512);
                __syncwarp();
                A = keys2;
                B = keys2 + 512;
                la = lb = 512;
            }
            const int perm = Epad >> 5;
            const int diag = perm * lane;
            int a0 = max(0, diag - lb), a1 = min(diag, la);
            while (a0 < a1) {
                const int mid = (a0 + a1) >> 1;
                if (A[mid] <= B[diag - 1 - mid])
                    a0 = mid + 1;
                else
                    a1 = mid;
            }
            int i = a0, j = diag - a0;
            uint32_t ka = (i < la) ? A[i] : 0xffffffffu, kb = (j < lb) ? B[j] : 0xffffffffu;
#pragma unroll 1
            for (int q = 0; q < perm; q++) {
                const bool takeA = ka <= kb;
                const uint32_t slot = (takeA ? ka : kb) & IDX_MASK;
                order[diag + q] = (uint16_t)slot;
                cnt_accumulate<NW>(local, (uint32_t)desc[slot], tab);
                if (takeA) {
                    i++;
                    ka = (i < la) ? A[i] : 0xffffffffu;
                } else {
                    j++;
                    kb = (j < lb) ? B[j] : 0xffffffffu;
                }
            }
            __syncwarp();
        }

        // ---- scan over the order list; lane owns positions [per*lane, per*lane+per)
        const int per = Epad >> 5;
        const int p0 = per * lane;
        bool recount = false;                // only after a repair pass changed the order list
        for (;;) {
            if (recount) {
                cnt_zero<NW>(local);
#pragma unroll 1
                for (int q = 0; q < per; q++)
                    cnt_accumulate<NW>(local, (uint32_t)desc[order[p0 + q]], tab);
            }
            double prevT = tm[order[lane == 0 ? 0 : p0 - 1]];
            int invAcc = 0;
            acc = 0.0;
            if constexpr (NW == 0) {
                uint32_t incl = local.w32;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(FULL_MASK, incl, o);
                    if (lane >= o)
                        incl += v;
                }
                const uint32_t b32 = incl - local.w32 + 0x40004000u;
                double k0 = (double)((int)(b32 & 0xffffu) - CNT_BIAS);
                double k1 = (double)((int)(b32 >> 16) - CNT_BIAS);
                const double i0 = tab->inv2N[0], i1 = tab->inv2N[1];
                const double b0 = tab->mu[0] - i0, b1 = tab->mu[1] - i1;
                double lambda = k0 * (k0 * i0 + b0) + k1 * (k1 * i1 + b1);
                if (lane == 0)       // event 0 is skipped (:101): its term is taken out again
                    acc = -reinterpret_cast<const ClsRec *>(clsBytes + desc[order[0]])->term;
#pragma unroll 2
                for (int q = 0; q < per; q++) {
                    const int slot = order[p0 + q];
                    const double x = tm[slot];
                    const ClsRec *rec = reinterpret_cast<const ClsRec *>(clsBytes + desc[slot]);
                    const double dt = x - prevT;
                    invAcc |= __double2hiint(dt);
                    acc = fma(-dt, lambda, acc);
                    prevT = x;
                    acc += rec->term;
                    k0 += __hiloint2double((int)rec->dk0hi, 0);
                    k1 += __hiloint2double((int)rec->dk1hi, 0);
                    lambda = k0 * (k0 * i0 + b0) + k1 * (k1 * i1 + b1);
                }
            } else {
                Counts<NW> base;
#pragma unroll
                for (int i = 0; i < NW; i++) {
                    unsigned long long incl = local.w[i];
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned long long v = shfl_up_u64(incl, o);
                        if (lane >= o)
                            incl += v;
                    }
                    base.w[i] = incl - local.w[i] + CNT_BIAS_WORD;
                }
                double lambda = 0.0;
#pragma unroll 1
                for (int c = 0; c < d; c++) {
                    const double kd = (double)cnt_get<NW>(base, c);
                    lambda += (kd * (kd - 1.0)) * tab->inv2N[c] + kd * tab->mu[c];
                }
                bool first = (lane == 0);
#pragma unroll 2
                for (int q = 0; q < per; q++) {
                    const int slot = order[p0 + q];
                    const double x = tm[slot];
                    const uint32_t ds = desc[slot];
                    const int kind = (ds >> 10) & 3, from = (ds >> 5) & 31, to = ds & 31;
                    const double dt = x - prevT;
                    invAcc |= __double2hiint(dt);
                    acc = fma(-dt, lambda, acc);
                    prevT = x;
                    const bool leave = (kind == KIND_COALESCE) | (kind == KIND_MIGRATE);
                    const bool enter = (kind == KIND_SAMPLE) | (kind == KIND_MIGRATE);
                    const int a = (kind == KIND_MIGRATE) ? to : from;
                    const double term = tab->logM[from * d + to];
                    acc += (leave && !first) ? term : 0.0;
                    const double kl = (double)cnt_get<NW>(base, from),
                                 ke = (double)cnt_get<NW>(base, a);
                    const double dl = (2.0 * (kl - 1.0)) * tab->inv2N[from] + tab->mu[from];
                    const double de = (2.0 * ke) * tab->inv2N[a] + tab->mu[a];
                    lambda += (enter ? de : 0.0) - (leave ? dl : 0.0);
                    cnt_add<NW>(base, from, -(long long)leave);
                    cnt_add<NW>(base, a, (long long)enter);
                    first = false;
                }
            }
            if (recount || !__any_sync(FULL_MASK, invAcc < 0))
                break;                                       // (NaN times cannot be ordered)
            // rare: quantisation ties out of exact order -> odd-even transposition on exact times
            bool swapped = true;
            while (swapped) {
                swapped = false;
#pragma unroll 1
                for (int parity = 0; parity < 2; parity++) {
#pragma unroll 1
                    for (int a = parity + 2 * lane; a + 1 < Epad; a += 64) {
                        const int ia = order[a], ib = order[a + 1];
                        if (tm[ia] > tm[ib]) {
                            order[a] = (uint16_t)ib;
                            order[a + 1] = (uint16_t)ia;
                            swapped = true;
                        }
                    }
                    __syncwarp();
                }
                swapped = __any_sync(FULL_MASK, swapped);
            }
            recount = true;
        }
        return selfMig ? -INFINITY : warp_sum_down(acc);
    }
    return NAN;      // not reached: the 256-slot stage only takes trees of the register path
}

// EMAX = 256 / 384 / 512 / 1024: every tree with n + M + 2 <= EMAX is evaluated here (3 CTAs =
// 24 warps per SM up to 512, 2 CTAs of 5 warps at 1024); larger ones are queued
// for the CTA-per-tree kernel, or -- when the caller promised a bound and no such kernel
// follows (p.queue == NULL) -- reported as NaN.  Persistent warps, one tree at a time per
// warp: lane 0 launches the bulk copies of the tree's array slices (TMA engine), the warp
// waits on its mbarrier, orders and scans.
// CTAs per SM of the stages up to 384 slots: shared memory allows four, but three (24 warps,
// 80 registers per thread) run as fast and leave the register path free of spills
#ifndef MTT_K1_MINB
#define MTT_K1_MINB 3
#endif
template <int NW, bool HAS_PARENT, bool PER_TREE_MODEL, int EMAX>
__global__ void __launch_bounds__(warps_per_cta(EMAX) * 32,
                                  (EMAX <= 384) ? MTT_K1_MINB : (EMAX <= 512 ? 3 : 2))
k1_density_warp(const K1Params p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ModelTab *tabs = reinterpret_cast<ModelTab *>(smem_raw);
    constexpr int W_WARPS = warps_per_cta(EMAX);
    constexpr int nTabs = PER_TREE_MODEL ? W_WARPS : 1;
    using WS = WarpSmem<NW, HAS_PARENT, EMAX>;
    WS *wsAll = reinterpret_cast<WS *>(smem_raw + sizeof(ModelTab) * nTabs);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WS *ws = wsAll + warp;
    ModelTab *tab = tabs + (PER_TREE_MODEL ? warp : 0);
    const int d = p.d;

    if (!PER_TREE_MODEL)
        fill_model_tab(tab, p.m, p.N, d, threadIdx.x, blockDim.x);
    if (lane == 0) {
        mbar_init((uint64_t *)&ws->bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    // (tree indices fit 32 bits: mtt_k1_launch refuses B >= 2^31)
    const int nWarpsTotal = (int)gridDim.x * W_WARPS, B = (int)p.B;
    uint32_t phase = 0;
    int t = (int)blockIdx.x * W_WARPS + warp;
    TreeSpan cur = load_span(p, t);
#ifdef MTT_K1_STAGGER       // experiment: spread the warps of an SM over the phases of a tree
    __nanosleep((unsigned)(MTT_K1_STAGGER * (warp + W_WARPS * (blockIdx.x % 3))));
#endif

#pragma unroll 1
    for (; t < B; t += nWarpsTotal) {
#ifdef MTT_K1_NO_SPAN_PREFETCH
        if (t != (int)blockIdx.x * W_WARPS + warp)
            cur = load_span(p, t);
        const TreeSpan nxt = cur;
#else
        // offsets of my next tree: the loads complete under this tree's work
        const TreeSpan nxt = load_span(p, (long long)t + nWarpsTotal);
#endif
        if (cur.n == 0) {            // no nodes: an empty sum
            if (lane == 0) {
                p.logP[t] = cur.M == 0 ? 0.0 : NAN;
                if (HAS_PARENT && p.valid)
                    p.valid[t] = cur.M == 0;
            }
        } else if (cur.n + cur.M + SLOT_SLACK > EMAX) {
            if (lane == 0) {
                if (p.queue) {       // too large for the warp path: queue for k1_density_cta
                    const int slot = atomicAdd(&p.queue[0], 1);
                    p.queue[1 + slot] = (int)t;
                } else {             // beyond the bound the caller gave: an error, loudly
                    p.logP[t] = NAN;
                    if (HAS_PARENT && p.valid)
                        p.valid[t] = 0;
                }
            }
        } else {
            const StageLayout L = make_layout<HAS_PARENT>(p, cur);
            if (lane == 0)
                issue_tree<HAS_PARENT, EMAX>(p, cur, L, &ws->st, (uint64_t *)&ws->bar);
#ifdef MTT_K1_LANE_STAGE
            stage_small<HAS_PARENT, EMAX>(p, cur, L, &ws->st, lane);
#endif
            if (PER_TREE_MODEL) {
                fill_model_tab(tab, p.m + (long long)t * d * d, p.N + (long long)t * d, d, lane, 32);
                __syncwarp();
            }
#ifdef MTT_K1_L2_PREFETCH     // bulk (TMA) prefetch: measured slower (405 vs 442 M trees/s): off
            if (lane == 0)
                prefetch_tree<HAS_PARENT>(p, nxt);       // (an empty span past the batch's end)
#endif
#ifdef MTT_K1_LINE_PREFETCH
            {   // one 128 B line of my next tree per lane into L2 (lanes 0-15 heights, 16-21
                // change times, 22-25 node types, 26-27 change types, 28-31 owners)
                const unsigned char *a;
                uint32_t bytes, i;
                if (lane < 16) {
                    a = (const unsigned char *)(p.height + nxt.n0); bytes = 8u * nxt.n; i = lane;
                } else if (lane < 22) {
                    a = (const unsigned char *)(p.chgTime + nxt.c0); bytes = 8u * nxt.M; i = lane - 16;
                } else if (lane < 26) {
                    a = (const unsigned char *)(p.nodeType + nxt.n0); bytes = nxt.n; i = lane - 22;
                } else if (lane < 28) {
                    a = (const unsigned char *)(p.chgType + nxt.c0); bytes = nxt.M; i = lane - 26;
                } else {
                    a = (const unsigned char *)(p.chgNode + nxt.c0); bytes = 2u * nxt.M; i = lane - 28;
                }
                if (128u * i < bytes + 127u && bytes)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a + 128u * i));
            }
#endif
            mbar_wait((uint64_t *)&ws->bar, phase);
            phase ^= 1u;
            __syncwarp();        // the lanes' own stores (stage_small) are visible to the warp
            bool ok = true, bad = false;
#ifdef MTT_K1_EXP_TWICE       // experiment: the work of two trees per staged tree
            (void)process_tree<NW, HAS_PARENT, EMAX>(ws, L, cur.n, cur.M, tab, d, lane, ok, bad);
            __syncwarp();
#endif
            const double acc = process_tree<NW, HAS_PARENT, EMAX>(ws, L, cur.n, cur.M, tab, d,
                                                                 lane, ok, bad);
            if (lane == 0) {
                if (HAS_PARENT && p.valid)
                    p.valid[t] = ok ? 1 : 0;
                // malformed input (owner / parent / type out of range, a change into its own
                // type): NaN; an invalid tree with checkValidity: -infinity (:90-91)
                p.logP[t] = bad ? NAN
                                : ((HAS_PARENT && p.checkValidity && !ok) ? -INFINITY : acc);
            }
            __syncwarp();      // the stage is free again
        }
        cur = nxt;
    }
}

// =====================================================================================
// CTA-per-tree kernel (queue consumer; also the direct path when every tree is large)
// =====================================================================================
// Trees beyond the warp stages (configs 3 and 4: 2.1 k and 10.4 k events), up to
// MTT_MAX_TREE_EVENTS.  The same plan as the warp path, spread over a CTA: S = the event count
// rounded up to whole runs; every warp orders runs of 256 slots in registers (warp_sort256),
// the runs are merged level by level with merge path (every thread produces 8 consecutive
// outputs per step), quantisation ties are put in exact order by odd-even passes over equal
// quanta, and each thread scans S / THREADS consecutive positions after a CTA-wide prefix sum
// of the packed lineage counts.  Slots without an event are twins of the last node, as above.
// Shared memory: two key buffers, the descriptors and -- when it fits -- the event times
// (18 bytes per slot, else 10 and the times are read through L1 / L2).
struct CtaQuant {
    double lo, scale;
    int shift;                         // slot bits of the key
};

template <int NW, int THREADS>
__global__ void __launch_bounds__(THREADS)
k1_density_cta(const K1Params p, int Scap, int tmInSmem, int useQueue, int minS, int nanBeyond)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ModelTab *tab = reinterpret_cast<ModelTab *>(smem_raw);
    unsigned char *sm = smem_raw + sizeof(ModelTab);
    double *tmS = nullptr;
    if (tmInSmem) {
        tmS = reinterpret_cast<double *>(sm);
        sm += (size_t)Scap * 8;
    }
    uint32_t *keysA = reinterpret_cast<uint32_t *>(sm);
    uint32_t *keysB = keysA + Scap;
    uint16_t *desc = reinterpret_cast<uint16_t *>(keysB + Scap);
    constexpr int NWARPS = THREADS / 32;
    __shared__ double redD[NWARPS];
    __shared__ unsigned long long redC[NWARPS][NW];
    __shared__ int redI[3 * NWARPS];
    __shared__ int sFlag;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = p.d;
    const long long nWork = useQueue ? (long long)p.queue[0] : p.B;

    for (long long wi = blockIdx.x; wi < nWork; wi += gridDim.x) {
        const long long t = useQueue ? (long long)p.queue[1 + wi] : wi;
        const long long n0 = p.nodeOff[t], c0 = p.chgBase[t];
        const int n = (int)(p.nodeOff[t + 1] - n0), M = (int)(p.chgBase[t + 1] - c0);
        const int E = n + M;
        const int nLeaves = (n + 1) >> 1;
        const double *H = p.height + n0;
        const double *CT = p.chgTime + c0;
        const uint8_t *T = p.nodeType + n0, *CTy = p.chgType + c0;
        const uint16_t *O = p.chgNode + c0;         // change owners
        const int *P = p.parent ? p.parent + n0 : nullptr;
        constexpr int RUN = 256 > THREADS ? 256 : THREADS;     // S is a multiple of both
        const int S = (E + RUN - 1) / RUN * RUN;

        __syncthreads();                            // the previous tree is done with smem
        if ((minS > 0 && S <= minS) || (S > Scap && !nanBeyond))
            continue;               // another launch of this call takes this size class
        if (S > Scap || E == 0) {   // beyond the capacity: NaN (host maps it to an error)
            if (tid == 0) {
                p.logP[t] = (E == 0) ? 0.0 : NAN;
                if (p.valid)
                    p.valid[t] = E == 0;
            }
            continue;
        }
        const bool perTree = p.nModels > 1;
        fill_model_tab(tab, p.m + (perTree ? t * d * d : 0), p.N + (perTree ? t * d : 0), d, tid,
                       THREADS);

        // ---- descriptors, staged times, validity, time range (high words)
        bool ok = true, bad = false, selfMig = false;
        HiWordRange hr;
        hr.init();
        const double t0 = H[n - 1];                 // twins copy the last node (see above)
        uint8_t *hasChg = reinterpret_cast<uint8_t *>(keysB);         // n bytes, validity only
        for (int e = tid; e < n; e += THREADS) {
            const double x = H[e];
            const int type = T[e];
            hr.add(x);
            if (tmS)
                tmS[e] = x;
            bad |= type >= d;
            desc[e] = (uint16_t)((e < nLeaves) ? (KIND_SAMPLE << 10 | (type & 31) << 5)
                                               : (KIND_COALESCE << 10 | (type & 31) << 5 |
                                                  (type & 31)));
            if (P)
                hasChg[e] = 0;
        }
        for (int s2 = E + tid; s2 < S; s2 += THREADS) {               // twins of the last node
            if (tmS)
                tmS[s2] = t0;
            desc[s2] = (uint16_t)(KIND_NONE << 10);
        }
        __syncthreads();
        for (int c = tid; c < M; c += THREADS) {
            const int node = O[c];
            const bool oob = node >= n;
            const int nd = oob ? 0 : node;
            const bool first = (c == 0) || ((int)O[c - 1] != node);
            const bool last = (c == M - 1) || ((int)O[c + 1] != node);
            const int par = P ? P[nd] : 0;
            const bool isRoot = P && par < 0;
            const int from = first ? (int)T[nd] : (int)CTy[c - 1];
            const int to = CTy[c];
            const double tc = CT[c];
            bad |= (from >= d) | (to >= d) | (P && par >= n);
            ok = ok && !oob;
            // ignored (:161-162), owner-less or self changes do nothing; the first two also
            // leave the time axis (twins of node 0)
            const bool gone = oob || isRoot;
            selfMig |= (from == to) && !gone;
            desc[n + c] = (gone || from == to)
                              ? (uint16_t)(KIND_NONE << 10)
                              : (uint16_t)(KIND_MIGRATE << 10 | (from & 31) << 5 | (to & 31));
            if (tmS)
                tmS[n + c] = (gone || from == to) ? t0 : tc;
            if (!gone && from != to)
                hr.add(tc);
            if (P) {
                const double before = first ? H[nd] : CT[c - 1];
                ok = ok && !(before > tc);
                if (last && !isRoot && !oob && par < n) {
                    ok = ok && !(tc > H[par]);
                    ok = ok && (to == (int)T[par]);
                }
                if (first && !oob)
                    hasChg[nd] = 1;
            }
        }
        if (P) {
            __syncthreads();
            for (int e = tid; e < n; e += THREADS) {
                const int par = P[e];
                bad |= par >= n;
                if (par >= 0 && par < n && !hasChg[e]) {
                    ok = ok && !(H[e] > H[par]);
                    ok = ok && (T[e] == T[par]);
                }
            }
        }
        // event time by slot: staged, or through the caches (a change that does nothing is a
        // twin of the last node, like the padding)
        auto ev_time = [&](int slot) -> double {
            if (tmS)
                return tmS[slot];
            if (slot < n)
                return H[slot];
            if (slot >= E || ((desc[slot] >> 10) & 3) == KIND_NONE)
                return t0;
            return CT[slot - n];
        };
        // CTA-wide reductions of the range and the verdicts
        {
            const int smin = __reduce_min_sync(FULL_MASK, hr.smin);
            const int smax = __reduce_max_sync(FULL_MASK, hr.smax);
            const uint32_t umax = __reduce_max_sync(FULL_MASK, hr.umax);
            const uint32_t f = __reduce_or_sync(FULL_MASK, (bad ? 1u : 0u) | (ok ? 0u : 2u) |
                                                               (selfMig ? 4u : 0u));
            if (tid == 0)
                sFlag = 0;
            __syncthreads();
            if (lane == 0) {
                redI[warp] = smin;
                redI[NWARPS + warp] = smax;
                redI[2 * NWARPS + warp] = (int)umax;
                atomicOr(&sFlag, (int)f);
            }
            __syncthreads();
            hr.init();
            for (int w = 0; w < NWARPS; w++) {
                hr.smin = min(hr.smin, redI[w]);
                hr.smax = max(hr.smax, redI[NWARPS + w]);
                hr.umax = max(hr.umax, (uint32_t)redI[2 * NWARPS + w]);
            }
            bad = sFlag & 1;
            ok = !(sFlag & 3);
            selfMig = sFlag & 4;
        }
        // Quantisation: slot bits = ceil(log2 S), the rest of the key is time.  With 10^4
        // events only 18 bits are left, and coalescent times crowd just after every sampling
        // date (thousands of lineages -> spacings of 1e-7): a uniform grid would put dozens of
        // events into one quantum.  The grid is equalised instead: a histogram of the times
        // over KB uniform bins, and every bin gets a share of the quanta proportional to its
        // events (at least one), i.e. the key is a piecewise-linear, monotone map of time that
        // is steep where the events are dense.
        CtaQuant qz;
        qz.shift = 32 - __clz(S - 1);
        const uint32_t slotMask = (1u << qz.shift) - 1u;
        const double QTOT = (double)((1u << (32 - qz.shift)) - 2u);
        int KB = 1 << (31 - __clz(S >> 2));
        KB = min(KB, 1024);
        uint32_t *hist = keysB, *qstart = keysB + KB;      // (keysB is free until the merges)
        {
            const int hwLo = ((int)hr.umax < 0) ? (int)hr.umax : hr.smin;
            const int hwHi = (hr.smax >= 0) ? hr.smax : hr.smin;
            qz.lo = __hiloint2double(hwLo, hwLo < 0 ? -1 : 0);
            const double hi = __hiloint2double(hwHi, hwHi < 0 ? 0 : -1);
            const double range = hi - qz.lo;
            qz.scale = (range > 1e-300 && range < INFINITY)
                           ? (double)KB * (1.0 - 1.0 / 1048576.0) / range : 0.0;   // bins per time
        }
        for (int i = tid; i < KB; i += THREADS)
            hist[i] = 0u;
        __syncthreads();
        for (int slot = tid; slot < S; slot += THREADS) {
            const int b = min(KB - 1, max(0, (int)((ev_time(slot) - qz.lo) * qz.scale)));
            atomicAdd(&hist[b], 1u);
        }
        __syncthreads();
        if (warp == 0) {                                   // exclusive prefix of the bins' quanta
            uint32_t carry = 0;
            for (int b0 = 0; b0 < KB; b0 += 32) {
                const uint32_t qn = 1u + (uint32_t)((QTOT - (double)KB) * (double)hist[b0 + lane] /
                                                    (double)S);
                uint32_t incl = qn;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(FULL_MASK, incl, o);
                    if (lane >= o)
                        incl += v;
                }
                qstart[b0 + lane] = carry + incl - qn;
                carry += __shfl_sync(FULL_MASK, incl, 31);
            }
            if (lane == 0)
                qstart[KB] = carry;
        }
        __syncthreads();
        auto quant = [&](double x) -> uint32_t {
            const double f = (x - qz.lo) * qz.scale;
            const int b = min(KB - 1, max(0, (int)f));
            const uint32_t q0 = qstart[b], qn = qstart[b + 1] - q0;
            const uint32_t in = (uint32_t)max(0.0, (f - (double)b) * (double)qn);
            return q0 + min(qn - 1u, in);
        };

        // ---- runs of 256 slots, ordered in registers
        for (int run = warp; run < S / 256; run += NWARPS) {
            uint32_t key[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int slot = 256 * run + lane + 32 * k;
                key[k] = (quant(ev_time(slot)) << qz.shift) | (uint32_t)slot;
            }
            warp_sort256(key, lane);
            uint4 *dst = reinterpret_cast<uint4 *>(keysA + 256 * run + 8 * lane);
            dst[0] = make_uint4(key[0], key[1], key[2], key[3]);
            dst[1] = make_uint4(key[4], key[5], key[6], key[7]);
        }
        __syncthreads();

        // ---- merge levels (merge path): every thread writes 8 consecutive outputs per step
        uint32_t *src = keysA, *dst = keysB;
        for (int L = 256; L < S; L <<= 1) {
            for (int o = tid * 8; o < S; o += THREADS * 8) {
                const int base = o & ~(2 * L - 1);
                const int la = min(L, S - base), lb = max(0, min(L, S - base - L));
                const uint32_t *A = src + base, *B = src + base + L;
                const int diag = o - base;
                int a0 = max(0, diag - lb), a1 = min(diag, la);
                while (a0 < a1) {
                    const int mid = (a0 + a1) >> 1;
                    if (A[mid] <= B[diag - 1 - mid])
                        a0 = mid + 1;
                    else
                        a1 = mid;
                }
                int i = a0, j = diag - a0;
                uint32_t ka = (i < la) ? A[i] : 0xffffffffu, kb = (j < lb) ? B[j] : 0xffffffffu;
                uint32_t out[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const bool takeA = ka <= kb;
                    out[q] = takeA ? ka : kb;
                    if (takeA) {
                        i++;
                        ka = (i < la) ? A[i] : 0xffffffffu;
                    } else {
                        j++;
                        kb = (j < lb) ? B[j] : 0xffffffffu;
                    }
                }
                uint4 *w = reinterpret_cast<uint4 *>(dst + o);
                w[0] = make_uint4(out[0], out[1], out[2], out[3]);
                w[1] = make_uint4(out[4], out[5], out[6], out[7]);
            }
            __syncthreads();
            uint32_t *tmp = src;
            src = dst;
            dst = tmp;
        }

        // ---- exact order inside equal quanta: odd-even passes until nothing moves
        for (;;) {
            bool moved = false;
            __syncthreads();
            if (tid == 0)
                sFlag = 0;
            for (int parity = 0; parity < 2; parity++) {
                __syncthreads();
                for (int a = 2 * tid + parity; a + 1 < S; a += 2 * THREADS) {
                    const uint32_t ka = src[a], kb = src[a + 1];
                    if ((ka >> qz.shift) == (kb >> qz.shift) &&
                        ev_time((int)(ka & slotMask)) > ev_time((int)(kb & slotMask))) {
                        src[a] = kb;
                        src[a + 1] = ka;
                        moved = true;
                    }
                }
            }
            if (moved)
                sFlag = 1;
            __syncthreads();
            if (!sFlag)
                break;
        }

        // ---- scan: thread owns sorted positions [q0, q0 + chunk)
        const int chunk = S / THREADS;
        const int q0 = tid * chunk;
        Counts<NW> local;
#pragma unroll
        for (int i = 0; i < NW; i++)
            local.w[i] = 0;
        for (int q = q0; q < q0 + chunk; q++)
            cnt_apply<NW>(local, desc[src[q] & slotMask]);
        Counts<NW> base;
#pragma unroll
        for (int i = 0; i < NW; i++) {
            unsigned long long incl2 = local.w[i];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long v = shfl_up_u64(incl2, o);
                if (lane >= o)
                    incl2 += v;
            }
            if (lane == 31)
                redC[warp][i] = incl2;
            base.w[i] = incl2 - local.w[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < NW; i++) {
            for (int w = 0; w < warp; w++)
                base.w[i] += redC[w][i];
            base.w[i] += CNT_BIAS_WORD;              // see CNT_BIAS: equal-time transients
        }
        double lambda = 0.0;
        for (int c = 0; c < d; c++) {
            const double kd = (double)cnt_get<NW>(base, c);
            lambda += (kd * (kd - 1.0)) * tab->inv2N[c] + kd * tab->mu[c];
        }
        double prevT = ev_time((int)(src[q0 > 0 ? q0 - 1 : 0] & slotMask));
        double acc = 0.0;
        bool first = (tid == 0);                     // event 0 is skipped (:101)
        for (int q = q0; q < q0 + chunk; q++) {
            const int slot = (int)(src[q] & slotMask);
            const double x = ev_time(slot);
            const uint32_t ds = desc[slot];
            const int kind = (ds >> 10) & 3, from = (ds >> 5) & 31, to = ds & 31;
            const double dt = x - prevT;
            acc = fma(-dt, lambda, acc);
            prevT = x;
            const bool leave = (kind == KIND_COALESCE) | (kind == KIND_MIGRATE);
            const bool enter = (kind == KIND_SAMPLE) | (kind == KIND_MIGRATE);
            const int a = (kind == KIND_MIGRATE) ? to : from;
            const double term = tab->logM[from * d + to];   // log(1/N) or log(m)  (:127-137)
            acc += (leave && !first) ? term : 0.0;
            const double kl = (double)cnt_get<NW>(base, from), ke = (double)cnt_get<NW>(base, a);
            const double dl = (2.0 * (kl - 1.0)) * tab->inv2N[from] + tab->mu[from];
            const double de = (2.0 * ke) * tab->inv2N[a] + tab->mu[a];
            lambda += (enter ? de : 0.0) - (leave ? dl : 0.0);
            cnt_add<NW>(base, from, -(long long)leave);
            cnt_add<NW>(base, a, (long long)enter);
            first = false;
        }
        acc = warp_sum_down(acc);
        __syncthreads();
        if (lane == 0)
            redD[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int w = 0; w < NWARPS; w++)
                tot += redD[w];
            if (P && p.valid)
                p.valid[t] = ok ? 1 : 0;
            p.logP[t] = bad ? NAN
                            : ((P && p.checkValidity && !ok) ? -INFINITY
                                                             : (selfMig ? -INFINITY : tot));
        }
    }
}

template <int NW, bool HP, bool PT, int EMAX>
int launch_warp(mtt_ctx *ctx, const K1Params &p)
{
    constexpr int W = warps_per_cta(EMAX);
    size_t smem = sizeof(ModelTab) * (PT ? W : 1) + sizeof(WarpSmem<NW, HP, EMAX>) * W;
    auto kern = k1_density_warp<NW, HP, PT, EMAX>;
    MTT_CUDA(ctx, mtt_fn_smem(ctx, kern, smem));
    int perSM = 1;
    MTT_CUDA(ctx, mtt_fn_occupancy(ctx, kern, W * 32, smem, &perSM));
    if (perSM < 1)
        perSM = 1;
    long long want = (p.B + W - 1) / W;
    long long grid = (long long)ctx->numSMs * perSM;
    if (want < grid)
        grid = want;
    kern<<<(unsigned)grid, W * 32, smem, ctx->stream>>>(p);
    MTT_CHECK_LAUNCH(ctx);
    return MTT_OK;
}

template <int NW, int EMAX>
int launch_warp_any(mtt_ctx *ctx, const K1Params &p)
{
    const bool hp = p.parent != nullptr, pt = p.nModels > 1;
    if (hp)
        return pt ? launch_warp<NW, true, true, EMAX>(ctx, p)
                  : launch_warp<NW, true, false, EMAX>(ctx, p);
    return pt ? launch_warp<NW, false, true, EMAX>(ctx, p)
              : launch_warp<NW, false, false, EMAX>(ctx, p);
}

template <int NW, int THREADS>
int launch_cta_t(mtt_ctx *ctx, const K1Params &p, int Scap, int useQueue, long long nWorkHint,
                 int minS, int nanBeyond)
{
    // 10 bytes per slot for the keys and descriptors, 8 more when the times are staged too
    const size_t fixed = sizeof(ModelTab) + 16;
    int tmInSmem = fixed + (size_t)Scap * 18 <= 200 * 1024;
    size_t smem = fixed + (size_t)Scap * (tmInSmem ? 18 : 10);
    auto kern = k1_density_cta<(NW == 0) ? 1 : NW, THREADS>;
    MTT_CUDA(ctx, mtt_fn_smem(ctx, kern, smem));
    int perSM = 1;
    MTT_CUDA(ctx, mtt_fn_occupancy(ctx, kern, THREADS, smem, &perSM));
    if (perSM < 1)
        perSM = 1;
    long long grid = (long long)ctx->numSMs * perSM;
    if (nWorkHint > 0 && nWorkHint < grid)
        grid = nWorkHint;
    kern<<<(unsigned)grid, THREADS, smem, ctx->stream>>>(p, Scap, tmInSmem, useQueue, minS, nanBeyond);
    MTT_CHECK_LAUNCH(ctx);
    return MTT_OK;
}

// Ecap: upper bound on the events of any tree the kernel will see.  Without one
// (Ecap = MTT_MAX_TREE_EVENTS) the queue is served by three launches, one per size class, so
// that mid-size trees do not pay for the shared memory of the largest possible one.
template <int NW>
int launch_block(mtt_ctx *ctx, const K1Params &p, int Ecap, int useQueue, long long nWorkHint)
{
    // 256 threads up to 4096 slots (several CTAs per SM), 512 beyond (one or two CTAs per SM)
    if (Ecap < MTT_MAX_TREE_EVENTS) {
        if (Ecap <= 4096)
            return launch_cta_t<NW, 256>(ctx, p, (Ecap + 255) / 256 * 256, useQueue, nWorkHint, 0, 1);
        return launch_cta_t<NW, 512>(ctx, p, (Ecap + 511) / 512 * 512, useQueue, nWorkHint, 0, 1);
    }
    int rc = launch_cta_t<NW, 256>(ctx, p, 2304, useQueue, nWorkHint, 0, 0);
    if (rc == MTT_OK)
        rc = launch_cta_t<NW, 256>(ctx, p, 4096, useQueue, nWorkHint, 2304, 0);
    if (rc == MTT_OK)
        rc = launch_cta_t<NW, 512>(ctx, p, MTT_MAX_TREE_EVENTS, useQueue, nWorkHint, 4096, 1);
    return rc;
}

// maxEvents: > 0 upper bound on nodes + changes of any tree, 0 unknown
#ifdef MTT_K1_PROBE      // development only: compile just the headline instantiation
template <int NW>
int launch_all(mtt_ctx *ctx, const K1Params &p, long long)
{
    K1Params q = p;
    q.queue = nullptr;
    return launch_warp<0, false, false, 384>(ctx, q);
}
#else
template <int NW>
int launch_all(mtt_ctx *ctx, const K1Params &p, long long maxEvents)
{
    const int Ecap = (maxEvents > 0 && maxEvents < MTT_MAX_TREE_EVENTS) ? (int)maxEvents
                                                                       : MTT_MAX_TREE_EVENTS;
    // A stage of EMAX slots takes trees with nodes + changes + 2 <= EMAX.  With a bound that
    // fits a stage: ONE launch and no queue -- a tree beyond the promised bound is reported
    // as NaN (an error of the caller), never silently skipped.
    if (maxEvents > 0 && maxEvents + SLOT_SLACK <= 1024) {
        K1Params q = p;
        q.queue = nullptr;
        if (maxEvents + SLOT_SLACK <= 256)
            return launch_warp_any<NW, 256>(ctx, q);
        if (maxEvents + SLOT_SLACK <= 384)          // the smaller stage of config 2
            return launch_warp_any<NW, 384>(ctx, q);
        if (maxEvents + SLOT_SLACK <= 512)
            return launch_warp_any<NW, 512>(ctx, q);
        return launch_warp_any<NW, 1024>(ctx, q);   // four sorted runs, two merge levels
    }
    // some trees may exceed the warp path: they are queued and consumed by the block kernel;
    // the queue counter is zeroed again for the next call.  A known bound above 1024 means
    // mid-size and large trees: the 1024-event stage takes what it can.
    int rc = (maxEvents > 1024) ? launch_warp_any<NW, 1024>(ctx, p)
                                : launch_warp_any<NW, 512>(ctx, p);
    if (rc != MTT_OK)
        return rc;
    rc = launch_block<NW>(ctx, p, Ecap, 1, 0);
    if (rc != MTT_OK)
        return rc;
    MTT_CUDA(ctx, cudaMemsetAsync(p.queue, 0, sizeof(int), ctx->stream));
    return MTT_OK;
}

#endif

// Per-node change counts -> per-change owner list (the layout the kernels read), one warp per
// tree.  Counts that do not add up to the tree's number of changes leave 0xffff owners behind,
// which the density kernels treat as a malformed tree.
__global__ void k1_expand_counts(long long B, const long long *nodeOff, const uint16_t *chgCount,
                                 const long long *chgBase, uint16_t *chgNode)
{
    const long long t = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= B)
        return;
    const long long n0 = nodeOff[t], c0 = chgBase[t];
    const int n = (int)(nodeOff[t + 1] - n0), M = (int)(chgBase[t + 1] - c0);
    const uint16_t *K = chgCount + n0;
    uint16_t *O = chgNode + c0;
    for (int c = lane; c < M; c += 32)
        O[c] = 0xffff;
    __syncwarp();
    int carry = 0;
    for (int base = 0; base < n; base += 32) {
        const int e = base + lane;
        const int k = (e < n) ? (int)K[e] : 0;
        int incl = k;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o)
                incl += v;
        }
        const int off = carry + incl - k;
        for (int j = 0; j < k && off + j < M; j++)
            O[off + j] = (uint16_t)e;
        carry += __shfl_sync(FULL_MASK, incl, 31);
    }
}

}  // namespace

int mtt_k1_launch(mtt_ctx *ctx, int64_t B, int d, int64_t nModels, const double *m,
                  const double *N, const int64_t *nodeOff, const double *height,
                  const int32_t *parent, const uint8_t *nodeType, const uint16_t *chgNode,
                  const int64_t *chgBase, const double *chgTime, const uint8_t *chgType,
                  int checkValidity, double *logP, uint8_t *valid, long long maxEvents)
{
    if (B <= 0)
        return MTT_OK;
    if (d < 1 || d > MTT_MAX_DEMES)
        return mtt_fail(ctx, MTT_ERR_TOO_LARGE, "density: d=%d outside [1,%d]", d, MTT_MAX_DEMES);
    if (nModels != 1 && nModels != B)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "density: nModels must be 1 or B");
    if (B > 0x7fffffffLL)
        return mtt_fail(ctx, MTT_ERR_TOO_LARGE, "density: B exceeds 2^31-1 trees per call");
    K1Params p;
    p.B = B; p.d = d; p.nModels = nModels; p.m = m; p.N = N;
    p.nodeOff = (const long long *)nodeOff; p.height = height; p.parent = parent;
    p.nodeType = nodeType; p.chgNode = chgNode; p.chgBase = (const long long *)chgBase;
    p.chgTime = chgTime; p.chgType = chgType; p.checkValidity = checkValidity;
    p.logP = logP; p.valid = valid;
    // work queue for the block kernel: [0] = count (zero between calls), [1..] = tree indices
    const size_t need = sizeof(int) * (size_t)(B + 1);
    const bool fresh = need + 256 > ctx->bufs[BUF_D_QUEUE].cap;
    p.queue = (int *)mtt_buf(ctx, BUF_D_QUEUE, need);
    if (!p.queue)
        return MTT_ERR_CUDA;
    if (fresh)
        MTT_CUDA(ctx, cudaMemsetAsync(p.queue, 0, sizeof(int), ctx->stream));
    if (d <= 2)
        return launch_all<0>(ctx, p, maxEvents);     // two-deme specialisation of the warp path
    if (d <= 4)
        return launch_all<1>(ctx, p, maxEvents);
    if (d <= 8)
        return launch_all<2>(ctx, p, maxEvents);
    return launch_all<4>(ctx, p, maxEvents);
}

int mtt_k1_expand_counts(mtt_ctx *ctx, int64_t B, const int64_t *nodeOff,
                         const uint16_t *chgCount, const int64_t *chgBase, uint16_t *chgNode)
{
    if (B <= 0)
        return MTT_OK;
    k1_expand_counts<<<(unsigned)((B + 7) / 8), 256, 0, ctx->stream>>>(
        B, (const long long *)nodeOff, chgCount, (const long long *)chgBase, chgNode);
    MTT_CHECK_LAUNCH(ctx);
    return MTT_OK;
}
