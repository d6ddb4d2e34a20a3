// mtt_common.cuh -- shared host/device plumbing for the sm_100a kernels behind include/mtt_b200.h
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/mtt_b200.h"

// ------------------------------------------------------------------ context

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct mtt_ctx {
    int device = 0;
    int numSMs = 148;
    cudaStream_t ownStream = nullptr;
    cudaStream_t stream = nullptr;   // stream in use (own or caller's)
    std::string err;
    int64_t launches = 0;
    // grow-only device scratch, indexed by purpose
    std::vector<DevBuf> bufs;
    // piece 1: caller's bound on nodes+changes per tree for the _dev entry (0 = unknown)
    long long k1MaxEvents = 0;
    int tipCount = 0;                   // leaves installed by mtt_sc_set_tips (0 = none)
    // piece 3 state
    int unifModels = 0, unifD = 0, unifSym = 0;
    int unifPowCap = 0;                 // powers allocated per model (entries 0..cap-1)
    std::vector<int> unifPowSize;       // powers valid per model
    std::vector<int> unifSteady;        // steady flag per model
    std::vector<double> unifMu;         // host copy of mu per model
    // copy stream + events of the chunked host-buffer density entry (mtt_ctx.cu)
    cudaStream_t copyStream = nullptr;
    cudaStream_t copyStream2 = nullptr;   // device -> host leg of the chunked retype call
    std::vector<cudaEvent_t> events;
    // pinned staging of small calls (Xfer below)
    void *pinIn = nullptr, *pinOut = nullptr;
    size_t pinInCap = 0, pinOutCap = 0;
    // per kernel: the dynamic shared-memory limit already granted and the last occupancy answer
    // (mtt_fn_smem / mtt_fn_occupancy: no runtime query on the latency path of small calls)
    struct FnState {
        size_t smemSet = 0, occSmem = (size_t)-1;
        int occThreads = -1, perSM = 0;
    };
    std::map<const void *, FnState> fns;
};

enum BufId {
    BUF_D_MODEL = 0, BUF_D_N, BUF_D_NODEOFF, BUF_D_HEIGHT, BUF_D_PARENT, BUF_D_NODETYPE,
    BUF_D_CHGCOUNT, BUF_D_CHGBASE, BUF_D_CHGTIME, BUF_D_CHGTYPE, BUF_D_LOGP, BUF_D_VALID,
    BUF_D_QUEUE, BUF_D_CHGNODE, BUF_D_TIPH, BUF_D_TIPT, BUF_D_INTH, BUF_D_INTT,
    BUF_U_Q, BUF_U_LOGQ, BUF_U_R, BUF_U_MU, BUF_U_POW, BUF_U_LOGPOW, BUF_U_LOGN, BUF_U_META, BUF_U_IN0, BUF_U_IN1, BUF_U_IN2,
    BUF_U_IN3, BUF_U_IN4, BUF_U_IN5, BUF_U_IN6, BUF_U_IN7, BUF_U_IN8, BUF_U_OUT0, BUF_U_OUT1,
    BUF_U_OUT2, BUF_U_OUT3, BUF_U_OUT4, BUF_U_OUT5, BUF_U_OUT6, BUF_U_OUT7, BUF_U_PBA,
    BUF_U_LIN0, BUF_U_LIN1, BUF_U_LIN2, BUF_U_LIN3, BUF_U_LIN4, BUF_U_LIN5, BUF_U_LIN6, BUF_U_LIN7,
    BUF_U_LOUT0, BUF_U_LOUT1,
    BUF_P_MODEL, BUF_P_N, BUF_P_NODEOFF, BUF_P_HEIGHT, BUF_P_PARENT, BUF_P_LEAFTYPE,
    BUF_P_ORDEROFF, BUF_P_ORDER, BUF_P_U, BUF_P_UOFF, BUF_P_LOGW, BUF_P_STATUS, BUF_P_USED,
    BUF_P_LOGP, BUF_P_OUTTYPE, BUF_P_NCHG, BUF_P_CHGTIME, BUF_P_CHGTYPE, BUF_P_CHGNODE,
    BUF_C_OPS, BUF_C_OPOFF, BUF_C_ORDER, BUF_C_KAPPA, BUF_C_LOGL,
    BUF_STAGE_IN, BUF_STAGE_OUT,
    BUF_COUNT
};

int mtt_fail(mtt_ctx *ctx, int code, const char *fmt, ...);
// piece 3: make sure R^0..R^K exist for every installed model (k3_unif.cu)
int mtt_unif_ensure_powers(mtt_ctx *ctx, int K);
// grow-only device buffer with 256 B of slack past `bytes` (bulk copies over-read <= 15 B)
void *mtt_buf(mtt_ctx *ctx, int id, size_t bytes);

// Host <-> device transfer plan of one host-buffer call.  Small calls -- the latency-bound
// sweeps of a few dozen chains -- pack every input into ONE pinned staging buffer and every
// output into another: two copies per call instead of some fifteen from pageable memory.
// Large calls copy array by array (packing would cost a pass over the host data).
class Xfer {
public:
    explicit Xfer(mtt_ctx *ctx, size_t stagedLimit = (size_t)1 << 20) : ctx_(ctx), limit_(stagedLimit) {}
    // declare an input / output array; *dptr receives its device address in prepare()
    template <typename T>
    void in(const T *src, size_t count, int bufId, T **dptr)
    {
        ins_.push_back(Item{(void *)src, count * sizeof(T), bufId, (void **)dptr, 0});
    }
    template <typename T>
    void out(T *dst, size_t count, int bufId, T **dptr)
    {
        outs_.push_back(Item{(void *)dst, count * sizeof(T), bufId, (void **)dptr, 0});
    }
    bool staged() const { return staged_; }
    // Staged outputs of at most this many bytes are written by the kernels straight into the
    // page-locked buffer (device-visible under unified addressing): no copy back, only the
    // synchronise.  For outputs that are written once and not read on the device.
    void direct_out(size_t limit = 4096) { directLimit_ = limit; }
    int prepare();      // device (and staging) memory, device addresses
    int upload();       // inputs -> device, asynchronous on the ctx stream
    int download();     // outputs -> host, synchronises the stream
    // staged mode: the host copy of an output after download_raw() (no scatter yet)
    template <typename T>
    const T *peek(const T *dst) const
    {
        for (const Item &o : outs_)
            if (o.host == (const void *)dst)
                return staged_ ? (const T *)((const char *)ctx_->pinOut + o.off) : nullptr;
        return nullptr;
    }

private:
    struct Item {
        void *host;
        size_t bytes;
        int bufId;
        void **dptr;
        size_t off;
    };
    mtt_ctx *ctx_;
    size_t limit_;
    bool staged_ = false, direct_ = false;
    size_t directLimit_ = 0;
    size_t inBytes_ = 0, outBytes_ = 0;
    std::vector<Item> ins_, outs_;
};

#define MTT_CUDA(ctx, expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess)                                                           \
            return mtt_fail((ctx), MTT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,         \
                            cudaGetErrorString(_e), __FILE__, __LINE__);                 \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per kernel and size (the limit only grows)
template <typename F>
inline cudaError_t mtt_fn_smem(mtt_ctx *ctx, F kernel, size_t smem)
{
    mtt_ctx::FnState &f = ctx->fns[(const void *)kernel];
    if (smem <= f.smemSet)
        return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute((const void *)kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
        f.smemSet = smem;
    return e;
}
// resident CTAs per SM, asked once per (kernel, threads, shared memory)
template <typename F>
inline cudaError_t mtt_fn_occupancy(mtt_ctx *ctx, F kernel, int threads, size_t smem, int *perSM)
{
    mtt_ctx::FnState &f = ctx->fns[(const void *)kernel];
    if (f.occThreads != threads || f.occSmem != smem) {
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, (const void *)kernel,
                                                                      threads, smem);
        if (e != cudaSuccess)
            return e;
        f.occThreads = threads;
        f.occSmem = smem;
        f.perSM = n;
    }
    *perSM = f.perSM;
    return cudaSuccess;
}

#define MTT_CHECK_LAUNCH(ctx)                                                            \
    do {                                                                                 \
        (ctx)->launches++;                                                               \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess)                                                           \
            return mtt_fail((ctx), MTT_ERR_CUDA, "kernel launch failed: %s (%s:%d)",     \
                            cudaGetErrorString(_e), __FILE__, __LINE__);                 \
    } while (0)

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__

#define FULL_MASK 0xffffffffu

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier + 1-D bulk async copy (TMA engine; SASS UBLKCP)
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
#ifdef MTT_MBAR_SUSPEND_HINT
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"   // suspend-time hint
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#endif
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy; src/dst 16 B aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes,
                                         uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// order-preserving map double -> uint64 (total order on non-NaN values)
__device__ __forceinline__ uint64_t f64_key(double x)
{
    uint64_t b = (uint64_t)__double_as_longlong(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
// fixed-order (lane 0 .. 31) reduction tree: deterministic run to run
__device__ __forceinline__ double warp_sum_down(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_down_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = fmin(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = fmax(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ uint64_t shfl_up_u64(uint64_t v, int o)
{
    return (uint64_t)__shfl_up_sync(FULL_MASK, (unsigned long long)v, o);
}

#endif  // __CUDACC__
