// mtt_ctx.cu -- context, error handling, device scratch, and the piece-1 C-ABI entry points.
#include "mtt_common.cuh"

#include <stdarg.h>

#include <algorithm>
#include <stdio.h>
#include <string.h>

static thread_local std::string g_create_err;

int mtt_fail(mtt_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx)
        ctx->err = buf;
    else
        g_create_err = buf;
    return code;
}

void *mtt_buf(mtt_ctx *ctx, int id, size_t bytes)
{
    DevBuf &b = ctx->bufs[id];
    if (bytes + 256 > b.cap) {
        if (b.p) {
            cudaStreamSynchronize(ctx->stream);
            cudaFree(b.p);
            b.p = nullptr;
            b.cap = 0;
        }
        size_t cap = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaMalloc(&b.p, cap);
        if (e != cudaSuccess) {
            mtt_fail(ctx, MTT_ERR_CUDA, "cudaMalloc(%zu) failed: %s", cap, cudaGetErrorString(e));
            b.p = nullptr;
            return nullptr;
        }
        b.cap = cap;
    }
    return b.p;
}

// ------------------------------------------------------------------ Xfer (mtt_common.cuh)
static int grow_pinned(mtt_ctx *ctx, void *&p, size_t &cap, size_t bytes)
{
    if (bytes <= cap)
        return MTT_OK;
    if (p) {
        cudaStreamSynchronize(ctx->stream);
        cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    const size_t want = bytes + bytes / 2 + 65536;
    MTT_CUDA(ctx, cudaMallocHost(&p, want));
    cap = want;
    return MTT_OK;
}

int Xfer::prepare()
{
    size_t in = 0, out = 0;
    for (Item &i : ins_) {
        i.off = in;
        in = (in + i.bytes + 255) & ~(size_t)255;
    }
    for (Item &o : outs_) {
        o.off = out;
        out = (out + o.bytes + 255) & ~(size_t)255;
    }
    inBytes_ = in;
    outBytes_ = out;
    staged_ = in + out <= limit_;
    if (staged_) {
        char *dIn = (char *)mtt_buf(ctx_, BUF_STAGE_IN, in);
        char *dOut = (char *)mtt_buf(ctx_, BUF_STAGE_OUT, out);
        if (!dIn || !dOut)
            return MTT_ERR_CUDA;
        int rc = grow_pinned(ctx_, ctx_->pinIn, ctx_->pinInCap, in);
        if (rc == MTT_OK)
            rc = grow_pinned(ctx_, ctx_->pinOut, ctx_->pinOutCap, out);
        if (rc != MTT_OK)
            return rc;
        for (Item &i : ins_)
            *i.dptr = dIn + i.off;
        direct_ = out > 0 && out <= directLimit_;
        for (Item &o : outs_)
            *o.dptr = (direct_ ? (char *)ctx_->pinOut : dOut) + o.off;
        return MTT_OK;
    }
    for (Item &i : ins_)
        if (!(*i.dptr = mtt_buf(ctx_, i.bufId, i.bytes)))
            return MTT_ERR_CUDA;
    for (Item &o : outs_)
        if (!(*o.dptr = mtt_buf(ctx_, o.bufId, o.bytes)))
            return MTT_ERR_CUDA;
    return MTT_OK;
}

int Xfer::upload()
{
    cudaStream_t s = ctx_->stream;
    if (staged_) {
        // the previous call's copy out of the staging buffer has completed (calls block)
        for (const Item &i : ins_)
            if (i.bytes)
                memcpy((char *)ctx_->pinIn + i.off, i.host, i.bytes);
        if (inBytes_)
            MTT_CUDA(ctx_, cudaMemcpyAsync(ctx_->bufs[BUF_STAGE_IN].p, ctx_->pinIn, inBytes_,
                                           cudaMemcpyHostToDevice, s));
        return MTT_OK;
    }
    for (const Item &i : ins_)
        if (i.bytes)
            MTT_CUDA(ctx_, cudaMemcpyAsync(*i.dptr, i.host, i.bytes, cudaMemcpyHostToDevice, s));
    return MTT_OK;
}

int Xfer::download()
{
    cudaStream_t s = ctx_->stream;
    if (staged_) {
        if (outBytes_ && !direct_)
            MTT_CUDA(ctx_, cudaMemcpyAsync(ctx_->pinOut, ctx_->bufs[BUF_STAGE_OUT].p, outBytes_,
                                           cudaMemcpyDeviceToHost, s));
        MTT_CUDA(ctx_, cudaStreamSynchronize(s));
        for (const Item &o : outs_)
            if (o.bytes)
                memcpy(o.host, (const char *)ctx_->pinOut + o.off, o.bytes);
        return MTT_OK;
    }
    for (const Item &o : outs_)
        if (o.bytes)
            MTT_CUDA(ctx_, cudaMemcpyAsync(o.host, *o.dptr, o.bytes, cudaMemcpyDeviceToHost, s));
    MTT_CUDA(ctx_, cudaStreamSynchronize(s));
    return MTT_OK;
}

extern "C" {

int mtt_abi_version(void)
{
    return MTT_ABI_VERSION;
}

int mtt_ctx_create(int device, mtt_ctx **out)
{
    if (!out)
        return mtt_fail(nullptr, MTT_ERR_INVALID_ARG, "mtt_ctx_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return mtt_fail(nullptr, MTT_ERR_NO_DEVICE,
                        "no CUDA device available (%s); this library has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0) {
        e = cudaGetDevice(&device);
        if (e != cudaSuccess)
            return mtt_fail(nullptr, MTT_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    }
    if (device >= count)
        return mtt_fail(nullptr, MTT_ERR_INVALID_ARG, "device %d out of range (%d devices)",
                        device, count);
    e = cudaSetDevice(device);
    if (e != cudaSuccess)
        return mtt_fail(nullptr, MTT_ERR_CUDA, "cudaSetDevice(%d): %s", device,
                        cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess)
        return mtt_fail(nullptr, MTT_ERR_CUDA, "cudaGetDeviceProperties: %s",
                        cudaGetErrorString(e));
    if (prop.major != 10)
        return mtt_fail(nullptr, MTT_ERR_NO_DEVICE,
                        "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                        device, prop.major, prop.minor);
    mtt_ctx *ctx = new mtt_ctx();
    ctx->device = device;
    ctx->numSMs = prop.multiProcessorCount;
    ctx->bufs.resize(BUF_COUNT);
    e = cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ctx;
        return mtt_fail(nullptr, MTT_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    ctx->stream = ctx->ownStream;
    *out = ctx;
    return MTT_OK;
}

int mtt_ctx_destroy(mtt_ctx *ctx)
{
    if (!ctx)
        return MTT_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &b : ctx->bufs)
        if (b.p)
            cudaFree(b.p);
    for (cudaEvent_t ev : ctx->events)
        cudaEventDestroy(ev);
    if (ctx->copyStream)
        cudaStreamDestroy(ctx->copyStream);
    if (ctx->copyStream2)
        cudaStreamDestroy(ctx->copyStream2);
    if (ctx->pinIn)
        cudaFreeHost(ctx->pinIn);
    if (ctx->pinOut)
        cudaFreeHost(ctx->pinOut);
    if (ctx->ownStream)
        cudaStreamDestroy(ctx->ownStream);
    delete ctx;
    return MTT_OK;
}

const char *mtt_last_error(const mtt_ctx *ctx)
{
    return ctx ? ctx->err.c_str() : g_create_err.c_str();
}

int mtt_ctx_set_stream(mtt_ctx *ctx, void *cuda_stream)
{
    if (!ctx)
        return MTT_ERR_INVALID_ARG;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->ownStream;
    return MTT_OK;
}

int mtt_ctx_synchronize(mtt_ctx *ctx)
{
    if (!ctx)
        return MTT_ERR_INVALID_ARG;
    MTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MTT_OK;
}

int64_t mtt_ctx_launch_count(const mtt_ctx *ctx)
{
    return ctx ? ctx->launches : 0;
}

int mtt_host_register(void *ptr, int64_t bytes)
{
    cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return mtt_fail(nullptr, MTT_ERR_CUDA, "cudaHostRegister: %s", cudaGetErrorString(e));
    }
    return MTT_OK;
}

int mtt_host_unregister(void *ptr)
{
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return mtt_fail(nullptr, MTT_ERR_CUDA, "cudaHostUnregister: %s", cudaGetErrorString(e));
    }
    return MTT_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ piece 1 entry points

int mtt_k1_launch(mtt_ctx *ctx, int64_t B, int d, int64_t nModels, const double *m,
                  const double *N, const int64_t *nodeOff, const double *height,
                  const int32_t *parent, const uint8_t *nodeType, const uint16_t *chgNode,
                  const int64_t *chgBase, const double *chgTime, const uint8_t *chgType,
                  int checkValidity, double *logP, uint8_t *valid, long long maxEvents);
int mtt_k1_expand_counts(mtt_ctx *ctx, int64_t B, const int64_t *nodeOff,
                         const uint16_t *chgCount, const int64_t *chgBase, uint16_t *chgNode);

// device-resident inputs; `owners` = chgNode layout, otherwise per-node counts (expanded here)
static int sc_logp_dev(mtt_ctx *ctx, const char *who, int64_t B, int d, int64_t nModels,
                       const double *m, const double *N, const int64_t *nodeOff,
                       const double *height, const int32_t *parent, const uint8_t *nodeType,
                       const uint16_t *chgCountOrNode, bool owners, const int64_t *chgBase,
                       const double *chgTime, const uint8_t *chgType, int checkValidity,
                       double *logP, uint8_t *valid)
{
    if (!ctx)
        return MTT_ERR_INVALID_ARG;
    if (B < 0 || !m || !N || !nodeOff || !height || !nodeType || !chgCountOrNode || !chgBase ||
        !logP)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "%s: NULL argument", who);
    if (B == 0)
        return MTT_OK;
    if (nModels != 1 && nModels != B)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "%s: nModels must be 1 or B", who);
    if (checkValidity && !parent)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "%s: checkValidity requires parent[]", who);
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint16_t *chgNode = chgCountOrNode;
    if (!owners) {
        // the kernels read the owner list: expand the counts (needs the number of changes)
        int64_t changes = 0;
        MTT_CUDA(ctx, cudaMemcpyAsync(&changes, chgBase + B, 8, cudaMemcpyDeviceToHost,
                                      ctx->stream));
        MTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        uint16_t *dO = (uint16_t *)mtt_buf(ctx, BUF_D_CHGNODE, (size_t)changes * 2);
        if (!dO)
            return MTT_ERR_CUDA;
        int rc = mtt_k1_expand_counts(ctx, B, nodeOff, chgCountOrNode, chgBase, dO);
        if (rc != MTT_OK)
            return rc;
        chgNode = dO;
    }
    return mtt_k1_launch(ctx, B, d, nModels, m, N, nodeOff, height, parent, nodeType, chgNode,
                         chgBase, chgTime, chgType, checkValidity, logP, valid, ctx->k1MaxEvents);
}

extern "C" int mtt_sc_logp_batch_dev(mtt_ctx *ctx, int64_t B, int d, int64_t nModels,
                                     const double *m, const double *N, const int64_t *nodeOff,
                                     const double *height, const int32_t *parent,
                                     const uint8_t *nodeType, const uint16_t *chgCount,
                                     const int64_t *chgBase, const double *chgTime,
                                     const uint8_t *chgType, int checkValidity, double *logP,
                                     uint8_t *valid)
{
    return sc_logp_dev(ctx, "mtt_sc_logp_batch_dev", B, d, nModels, m, N, nodeOff, height, parent,
                       nodeType, chgCount, false, chgBase, chgTime, chgType, checkValidity, logP,
                       valid);
}

extern "C" int mtt_sc_logp_owner_batch_dev(mtt_ctx *ctx, int64_t B, int d, int64_t nModels,
                                           const double *m, const double *N,
                                           const int64_t *nodeOff, const double *height,
                                           const int32_t *parent, const uint8_t *nodeType,
                                           const uint16_t *chgNode, const int64_t *chgBase,
                                           const double *chgTime, const uint8_t *chgType,
                                           int checkValidity, double *logP, uint8_t *valid)
{
    return sc_logp_dev(ctx, "mtt_sc_logp_owner_batch_dev", B, d, nModels, m, N, nodeOff, height,
                       parent, nodeType, chgNode, true, chgBase, chgTime, chgType, checkValidity,
                       logP, valid);
}

// host buffers
static int sc_logp_host(mtt_ctx *ctx, const char *who, int64_t B, int d, int64_t nModels,
                        const double *m, const double *N, const int64_t *nodeOff,
                        const double *height, const int32_t *parent, const uint8_t *nodeType,
                        const uint16_t *chgCountOrNode, bool owners, const int64_t *chgBase,
                        const double *chgTime, const uint8_t *chgType, int checkValidity,
                        double *logP, uint8_t *valid)
{
    if (!ctx)
        return MTT_ERR_INVALID_ARG;
    if (B < 0 || !m || !N || !nodeOff || !height || !nodeType || !chgBase || !logP)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "%s: NULL argument", who);
    if (B == 0)
        return MTT_OK;
    if (nModels != 1 && nModels != B)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "%s: nModels must be 1 or B", who);
    if (checkValidity && !parent)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "%s: checkValidity requires parent[]", who);
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));

    const int64_t nodes = nodeOff[B] - nodeOff[0], changes = chgBase[B] - chgBase[0];
    if (nodeOff[0] != 0 || chgBase[0] != 0)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "nodeOff[0] and chgBase[0] must be 0");
    if (changes > 0 && (!chgTime || !chgType || (owners && !chgCountOrNode)))
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "chgTime/chgType/chgNode NULL with changes present");
    if (!owners && !chgCountOrNode)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "%s: chgCount is NULL", who);
    long long maxEv = 0;
    for (int64_t b = 0; b < B; b++) {
        long long e = (nodeOff[b + 1] - nodeOff[b]) + (chgBase[b + 1] - chgBase[b]);
        if (nodeOff[b + 1] < nodeOff[b] || chgBase[b + 1] < chgBase[b])
            return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "offsets of tree %lld not monotone",
                            (long long)b);
        if (e > maxEv)
            maxEv = e;
    }
    if (maxEv > MTT_MAX_TREE_EVENTS)
        return mtt_fail(ctx, MTT_ERR_TOO_LARGE,
                        "a tree has %lld events (nodes + changes); limit is %d", maxEv,
                        MTT_MAX_TREE_EVENTS);

    cudaStream_t s = ctx->stream;
    const size_t md = (size_t)nModels * d * d * 8, nd = (size_t)nModels * d * 8;
    {
        // Small calls -- one tree of a BEAST chain, the few dozen trees of an ensemble step --
        // are latency: every input goes through one page-locked staging buffer and one copy,
        // the results come back in another (Xfer), instead of a dozen copies from pageable
        // memory.
        const size_t small = md + nd + (size_t)(B + 1) * 16 +
                             (size_t)nodes * (8 + 1 + (owners ? 0 : 2) + (parent ? 4 : 0)) +
                             (size_t)changes * (8 + 1 + 2) + (size_t)B * 9;
        if (small + 16 * 256 <= ((size_t)1 << 20)) {
            Xfer x(ctx);
            double *dM, *dN, *dH, *dCT = nullptr, *dLogP;
            long long *dNodeOff, *dChgBase;
            uint8_t *dT, *dCTy = nullptr, *dValid = nullptr;
            uint16_t *dK = nullptr, *dO = nullptr;
            int32_t *dP = nullptr;
            x.in(m, md / 8, BUF_D_MODEL, &dM);
            x.in(N, nd / 8, BUF_D_N, &dN);
            x.in((const long long *)nodeOff, (size_t)B + 1, BUF_D_NODEOFF, &dNodeOff);
            x.in((const long long *)chgBase, (size_t)B + 1, BUF_D_CHGBASE, &dChgBase);
            x.in(height, (size_t)nodes, BUF_D_HEIGHT, &dH);
            x.in(nodeType, (size_t)nodes, BUF_D_NODETYPE, &dT);
            if (!owners)
                x.in(chgCountOrNode, (size_t)nodes, BUF_D_CHGCOUNT, &dK);
            else if (changes > 0)
                x.in(chgCountOrNode, (size_t)changes, BUF_D_CHGNODE, &dO);
            if (changes > 0) {
                x.in(chgTime, (size_t)changes, BUF_D_CHGTIME, &dCT);
                x.in(chgType, (size_t)changes, BUF_D_CHGTYPE, &dCTy);
            }
            if (parent)
                x.in(parent, (size_t)nodes, BUF_D_PARENT, &dP);
            x.out(logP, (size_t)B, BUF_D_LOGP, &dLogP);
            if (valid && parent)
                x.out(valid, (size_t)B, BUF_D_VALID, &dValid);
            x.direct_out();                 // logP / valid: written once per tree
            int rc = x.prepare();
            if (rc == MTT_OK && x.staged()) {
                // (arrays that a tree set without changes never reads still need an address,
                // and the expanded owner list a place)
                if (!owners)
                    dO = (uint16_t *)mtt_buf(ctx, BUF_D_CHGNODE, (size_t)changes * 2);
                if (!dO)
                    dO = (uint16_t *)mtt_buf(ctx, BUF_D_CHGNODE, 16);
                if (!dCT)
                    dCT = (double *)mtt_buf(ctx, BUF_D_CHGTIME, 16);
                if (!dCTy)
                    dCTy = (uint8_t *)mtt_buf(ctx, BUF_D_CHGTYPE, 16);
                if (!dO || !dCT || !dCTy)
                    return MTT_ERR_CUDA;
                if ((rc = x.upload()) != MTT_OK)
                    return rc;
                if (!owners)
                    rc = mtt_k1_expand_counts(ctx, B, (const int64_t *)dNodeOff, dK,
                                              (const int64_t *)dChgBase, dO);
                if (rc != MTT_OK)
                    return rc;
                rc = mtt_k1_launch(ctx, B, d, nModels > 1 ? B : 1, dM, dN,
                                   (const int64_t *)dNodeOff, dH, dP, dT, dO,
                                   (const int64_t *)dChgBase, dCT, dCTy, checkValidity, dLogP,
                                   dValid, maxEv > 0 ? maxEv : 1);
                if (rc != MTT_OK)
                    return rc;
                if ((rc = x.download()) != MTT_OK)
                    return rc;
                if (valid && !parent)
                    memset(valid, 1, (size_t)B);
                return MTT_OK;
            }
            if (rc != MTT_OK)
                return rc;
        }
    }
    double *dM = (double *)mtt_buf(ctx, BUF_D_MODEL, md);
    double *dN = (double *)mtt_buf(ctx, BUF_D_N, nd);
    int64_t *dNodeOff = (int64_t *)mtt_buf(ctx, BUF_D_NODEOFF, (B + 1) * 8);
    int64_t *dChgBase = (int64_t *)mtt_buf(ctx, BUF_D_CHGBASE, (B + 1) * 8);
    double *dH = (double *)mtt_buf(ctx, BUF_D_HEIGHT, nodes * 8);
    uint8_t *dT = (uint8_t *)mtt_buf(ctx, BUF_D_NODETYPE, nodes);
    uint16_t *dK = owners ? nullptr : (uint16_t *)mtt_buf(ctx, BUF_D_CHGCOUNT, nodes * 2);
    uint16_t *dO = (uint16_t *)mtt_buf(ctx, BUF_D_CHGNODE, changes * 2);
    double *dCT = (double *)mtt_buf(ctx, BUF_D_CHGTIME, changes * 8);
    uint8_t *dCTy = (uint8_t *)mtt_buf(ctx, BUF_D_CHGTYPE, changes);
    int32_t *dP = parent ? (int32_t *)mtt_buf(ctx, BUF_D_PARENT, nodes * 4) : nullptr;
    double *dLogP = (double *)mtt_buf(ctx, BUF_D_LOGP, B * 8);
    uint8_t *dValid = (valid && parent) ? (uint8_t *)mtt_buf(ctx, BUF_D_VALID, B) : nullptr;
    if (!dM || !dN || !dNodeOff || !dChgBase || !dH || !dT || (!owners && !dK) || !dO || !dCT ||
        !dCTy || !dLogP || (parent && !dP) || (valid && parent && !dValid))
        return MTT_ERR_CUDA;

    MTT_CUDA(ctx, cudaMemcpyAsync(dM, m, md, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dN, N, nd, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dNodeOff, nodeOff, (B + 1) * 8, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dChgBase, chgBase, (B + 1) * 8, cudaMemcpyHostToDevice, s));

    // Large batches are cut into chunks of trees: the arrays of chunk k+1 cross the PCIe link
    // (copy stream) while chunk k is evaluated and its results go back (ctx stream).  The link
    // is the bottleneck of this entry; the pipeline hides everything else behind it.
    const size_t bytes = (size_t)nodes * (8 + 1 + (owners ? 0 : 2) + (parent ? 4 : 0)) +
                         (size_t)changes * (8 + 1 + (owners ? 2 : 0));
    const int nChunks = (bytes > ((size_t)32 << 20) && B >= 1024) ? 8 : 1;
    cudaStream_t cs = s;
    if (nChunks > 1) {
        if (!ctx->copyStream)
            MTT_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
        while ((int)ctx->events.size() < nChunks + 1) {
            cudaEvent_t ev;
            MTT_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            ctx->events.push_back(ev);
        }
        cs = ctx->copyStream;
        // the copy stream starts after whatever the ctx stream still has in flight on these
        // buffers (an earlier call's kernels) and after the offsets above
        MTT_CUDA(ctx, cudaEventRecord(ctx->events[nChunks], s));
        MTT_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->events[nChunks], 0));
    }
    int rc = MTT_OK;
    for (int k = 0; k < nChunks; k++) {
        const int64_t b0 = B * k / nChunks, b1 = B * (k + 1) / nChunks, Bk = b1 - b0;
        if (Bk == 0)
            continue;
        const int64_t n0 = nodeOff[b0], n1 = nodeOff[b1], c0 = chgBase[b0], c1 = chgBase[b1];
        const size_t nn = (size_t)(n1 - n0), nc = (size_t)(c1 - c0);
        MTT_CUDA(ctx, cudaMemcpyAsync(dH + n0, height + n0, nn * 8, cudaMemcpyHostToDevice, cs));
        MTT_CUDA(ctx, cudaMemcpyAsync(dT + n0, nodeType + n0, nn, cudaMemcpyHostToDevice, cs));
        if (!owners)
            MTT_CUDA(ctx, cudaMemcpyAsync(dK + n0, chgCountOrNode + n0, nn * 2,
                                          cudaMemcpyHostToDevice, cs));
        if (nc > 0) {
            MTT_CUDA(ctx, cudaMemcpyAsync(dCT + c0, chgTime + c0, nc * 8, cudaMemcpyHostToDevice, cs));
            MTT_CUDA(ctx, cudaMemcpyAsync(dCTy + c0, chgType + c0, nc, cudaMemcpyHostToDevice, cs));
            if (owners)
                MTT_CUDA(ctx, cudaMemcpyAsync(dO + c0, chgCountOrNode + c0, nc * 2,
                                              cudaMemcpyHostToDevice, cs));
        }
        if (parent)
            MTT_CUDA(ctx, cudaMemcpyAsync(dP + n0, parent + n0, nn * 4, cudaMemcpyHostToDevice, cs));
        if (nChunks > 1) {
            MTT_CUDA(ctx, cudaEventRecord(ctx->events[k], cs));
            MTT_CUDA(ctx, cudaStreamWaitEvent(s, ctx->events[k], 0));
        }
        if (!owners)
            rc = mtt_k1_expand_counts(ctx, Bk, dNodeOff + b0, dK, dChgBase + b0, dO);
        if (rc != MTT_OK)
            return rc;
        const bool perTree = nModels > 1;
        rc = mtt_k1_launch(ctx, Bk, d, perTree ? Bk : 1, dM + (perTree ? b0 * d * d : 0),
                           dN + (perTree ? b0 * d : 0), dNodeOff + b0, dH, dP, dT, dO,
                           dChgBase + b0, dCT, dCTy, checkValidity, dLogP + b0,
                           dValid ? dValid + b0 : nullptr, maxEv > 0 ? maxEv : 1);
        if (rc != MTT_OK)
            return rc;
        MTT_CUDA(ctx, cudaMemcpyAsync(logP + b0, dLogP + b0, Bk * 8, cudaMemcpyDeviceToHost, s));
        if (dValid)
            MTT_CUDA(ctx, cudaMemcpyAsync(valid + b0, dValid + b0, Bk, cudaMemcpyDeviceToHost, s));
    }
    MTT_CUDA(ctx, cudaStreamSynchronize(s));
    if (valid && !parent)
        memset(valid, 1, (size_t)B);
    return MTT_OK;
}

extern "C" int mtt_sc_logp_batch(mtt_ctx *ctx, int64_t B, int d, int64_t nModels,
                                 const double *m, const double *N, const int64_t *nodeOff,
                                 const double *height, const int32_t *parent,
                                 const uint8_t *nodeType, const uint16_t *chgCount,
                                 const int64_t *chgBase, const double *chgTime,
                                 const uint8_t *chgType, int checkValidity, double *logP,
                                 uint8_t *valid)
{
    return sc_logp_host(ctx, "mtt_sc_logp_batch", B, d, nModels, m, N, nodeOff, height, parent,
                        nodeType, chgCount, false, chgBase, chgTime, chgType, checkValidity, logP,
                        valid);
}

extern "C" int mtt_sc_logp_owner_batch(mtt_ctx *ctx, int64_t B, int d, int64_t nModels,
                                       const double *m, const double *N, const int64_t *nodeOff,
                                       const double *height, const int32_t *parent,
                                       const uint8_t *nodeType, const uint16_t *chgNode,
                                       const int64_t *chgBase, const double *chgTime,
                                       const uint8_t *chgType, int checkValidity, double *logP,
                                       uint8_t *valid)
{
    return sc_logp_host(ctx, "mtt_sc_logp_owner_batch", B, d, nModels, m, N, nodeOff, height,
                        parent, nodeType, chgNode, true, chgBase, chgTime, chgType, checkValidity,
                        logP, valid);
}

// ------------------------------------------------------------------ trees over one taxon set
// leaves from the installed tip arrays, internal nodes from the call's arrays -> full tree SoA
__global__ void k1_expand_tips(long long B, int nL, const double *leafH, const uint8_t *leafT,
                               const double *intH, const uint8_t *intT, long long t0,
                               double *height, uint8_t *nodeType, long long *nodeOff)
{
    const int nN = 2 * nL - 1, nI = nL - 1;
    const long long total = B * nN;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long t = i / nN;
        const int e = (int)(i - t * nN);
        const long long g = (t0 + t) * nN + e;
        if (e < nL) {
            height[g] = leafH[e];
            nodeType[g] = leafT[e];
        } else {
            height[g] = intH[(t0 + t) * nI + (e - nL)];
            nodeType[g] = intT[(t0 + t) * nI + (e - nL)];
        }
        if (e == 0)
            nodeOff[t0 + t] = (t0 + t) * nN;
        if (i == total - 1)
            nodeOff[t0 + B] = (t0 + B) * nN;
    }
}

extern "C" int mtt_sc_set_tips(mtt_ctx *ctx, int nLeaves, const double *leafHeight,
                               const uint8_t *leafType)
{
    if (!ctx)
        return MTT_ERR_INVALID_ARG;
    if (nLeaves < 2 || !leafHeight || !leafType)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "mtt_sc_set_tips: bad arguments");
    if (2 * nLeaves - 1 > 65535)
        return mtt_fail(ctx, MTT_ERR_TOO_LARGE, "mtt_sc_set_tips: more than 32768 tips");
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    double *dH = (double *)mtt_buf(ctx, BUF_D_TIPH, (size_t)nLeaves * 8);
    uint8_t *dT = (uint8_t *)mtt_buf(ctx, BUF_D_TIPT, (size_t)nLeaves);
    if (!dH || !dT)
        return MTT_ERR_CUDA;
    MTT_CUDA(ctx, cudaMemcpyAsync(dH, leafHeight, (size_t)nLeaves * 8, cudaMemcpyHostToDevice,
                                  ctx->stream));
    MTT_CUDA(ctx, cudaMemcpyAsync(dT, leafType, (size_t)nLeaves, cudaMemcpyHostToDevice,
                                  ctx->stream));
    MTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->tipCount = nLeaves;
    return MTT_OK;
}

extern "C" int mtt_sc_logp_tips_batch(mtt_ctx *ctx, int64_t B, int d, int64_t nModels,
                                      const double *m, const double *N, const double *intHeight,
                                      const int32_t *parent, const uint8_t *intType,
                                      const uint16_t *chgNode, const int64_t *chgBase,
                                      const double *chgTime, const uint8_t *chgType,
                                      int checkValidity, double *logP, uint8_t *valid)
{
    const char *who = "mtt_sc_logp_tips_batch";
    if (!ctx)
        return MTT_ERR_INVALID_ARG;
    if (ctx->tipCount < 2)
        return mtt_fail(ctx, MTT_ERR_STATE, "%s: mtt_sc_set_tips has not been called", who);
    if (B < 0 || !m || !N || !intHeight || !intType || !chgBase || !logP)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "%s: NULL argument", who);
    if (B == 0)
        return MTT_OK;
    if (nModels != 1 && nModels != B)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "%s: nModels must be 1 or B", who);
    if (checkValidity && !parent)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "%s: checkValidity requires parent[]", who);
    MTT_CUDA(ctx, cudaSetDevice(ctx->device));
    const int nL = ctx->tipCount, nN = 2 * nL - 1, nI = nL - 1;
    const int64_t nodes = B * nN, changes = chgBase[B] - chgBase[0];
    if (chgBase[0] != 0)
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "chgBase[0] must be 0");
    if (changes > 0 && (!chgTime || !chgType || !chgNode))
        return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "chgTime/chgType/chgNode NULL with changes present");
    long long maxEv = 0;
    for (int64_t b = 0; b < B; b++) {
        if (chgBase[b + 1] < chgBase[b])
            return mtt_fail(ctx, MTT_ERR_INVALID_ARG, "offsets of tree %lld not monotone", (long long)b);
        maxEv = std::max<long long>(maxEv, nN + (chgBase[b + 1] - chgBase[b]));
    }
    if (maxEv > MTT_MAX_TREE_EVENTS)
        return mtt_fail(ctx, MTT_ERR_TOO_LARGE, "a tree has %lld events (nodes + changes); limit is %d",
                        maxEv, MTT_MAX_TREE_EVENTS);
    cudaStream_t s = ctx->stream;
    const size_t md = (size_t)nModels * d * d * 8, nd = (size_t)nModels * d * 8;
    double *dM = (double *)mtt_buf(ctx, BUF_D_MODEL, md);
    double *dN = (double *)mtt_buf(ctx, BUF_D_N, nd);
    int64_t *dNodeOff = (int64_t *)mtt_buf(ctx, BUF_D_NODEOFF, (B + 1) * 8);
    int64_t *dChgBase = (int64_t *)mtt_buf(ctx, BUF_D_CHGBASE, (B + 1) * 8);
    double *dH = (double *)mtt_buf(ctx, BUF_D_HEIGHT, nodes * 8);
    uint8_t *dT = (uint8_t *)mtt_buf(ctx, BUF_D_NODETYPE, nodes);
    double *dIH = (double *)mtt_buf(ctx, BUF_D_INTH, (size_t)B * nI * 8);
    uint8_t *dIT = (uint8_t *)mtt_buf(ctx, BUF_D_INTT, (size_t)B * nI);
    uint16_t *dO = (uint16_t *)mtt_buf(ctx, BUF_D_CHGNODE, changes * 2);
    double *dCT = (double *)mtt_buf(ctx, BUF_D_CHGTIME, changes * 8);
    uint8_t *dCTy = (uint8_t *)mtt_buf(ctx, BUF_D_CHGTYPE, changes);
    int32_t *dP = parent ? (int32_t *)mtt_buf(ctx, BUF_D_PARENT, nodes * 4) : nullptr;
    double *dLogP = (double *)mtt_buf(ctx, BUF_D_LOGP, B * 8);
    uint8_t *dValid = (valid && parent) ? (uint8_t *)mtt_buf(ctx, BUF_D_VALID, B) : nullptr;
    if (!dM || !dN || !dNodeOff || !dChgBase || !dH || !dT || !dIH || !dIT || !dO || !dCT ||
        !dCTy || !dLogP || (parent && !dP) || (valid && parent && !dValid))
        return MTT_ERR_CUDA;
    const double *dTipH = (const double *)ctx->bufs[BUF_D_TIPH].p;
    const uint8_t *dTipT = (const uint8_t *)ctx->bufs[BUF_D_TIPT].p;
    const size_t bytes = (size_t)B * nI * 9 + (size_t)nodes * (parent ? 4 : 0) + (size_t)changes * 11;
    if (md + nd + (size_t)(B + 1) * 8 + bytes + (size_t)B * 9 + 12 * 256 <= ((size_t)1 << 20)) {
        // Small calls -- calculateLogP() of one chain, an ensemble step -- are latency: all
        // inputs through one page-locked staging buffer and one copy, the results written by
        // the kernel straight into page-locked memory (Xfer), as in mtt_sc_logp_owner_batch.
        Xfer x(ctx);
        x.in(m, md / 8, BUF_D_MODEL, &dM);
        x.in(N, nd / 8, BUF_D_N, &dN);
        x.in(chgBase, (size_t)B + 1, BUF_D_CHGBASE, &dChgBase);
        x.in(intHeight, (size_t)B * nI, BUF_D_INTH, &dIH);
        x.in(intType, (size_t)B * nI, BUF_D_INTT, &dIT);
        if (changes > 0) {
            x.in(chgTime, (size_t)changes, BUF_D_CHGTIME, &dCT);
            x.in(chgType, (size_t)changes, BUF_D_CHGTYPE, &dCTy);
            x.in(chgNode, (size_t)changes, BUF_D_CHGNODE, &dO);
        }
        if (parent)
            x.in(parent, (size_t)nodes, BUF_D_PARENT, &dP);
        x.out(logP, (size_t)B, BUF_D_LOGP, &dLogP);
        if (dValid)
            x.out(valid, (size_t)B, BUF_D_VALID, &dValid);
        x.direct_out();
        int rc = x.prepare();
        if (rc == MTT_OK)
            rc = x.upload();
        if (rc != MTT_OK)
            return rc;
        const long long total = (long long)B * nN;
        const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 32);
        k1_expand_tips<<<grid, 256, 0, s>>>(B, nL, dTipH, dTipT, dIH, dIT, 0, dH, dT,
                                            (long long *)dNodeOff);
        MTT_CHECK_LAUNCH(ctx);
        rc = mtt_k1_launch(ctx, B, d, nModels > 1 ? B : 1, dM, dN, dNodeOff, dH, dP, dT, dO,
                           dChgBase, dCT, dCTy, checkValidity, dLogP, dValid, maxEv);
        if (rc != MTT_OK)
            return rc;
        if ((rc = x.download()) != MTT_OK)
            return rc;
        if (valid && !parent)
            memset(valid, 1, (size_t)B);
        return MTT_OK;
    }
    MTT_CUDA(ctx, cudaMemcpyAsync(dM, m, md, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dN, N, nd, cudaMemcpyHostToDevice, s));
    MTT_CUDA(ctx, cudaMemcpyAsync(dChgBase, chgBase, (B + 1) * 8, cudaMemcpyHostToDevice, s));
    const int nChunks = (bytes > ((size_t)32 << 20) && B >= 1024) ? 8 : 1;
    cudaStream_t cs = s;
    if (nChunks > 1) {
        if (!ctx->copyStream)
            MTT_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
        while ((int)ctx->events.size() < nChunks + 1) {
            cudaEvent_t ev;
            MTT_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            ctx->events.push_back(ev);
        }
        cs = ctx->copyStream;
        MTT_CUDA(ctx, cudaEventRecord(ctx->events[nChunks], s));
        MTT_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->events[nChunks], 0));
    }
    for (int k = 0; k < nChunks; k++) {
        const int64_t b0 = B * k / nChunks, b1 = B * (k + 1) / nChunks, Bk = b1 - b0;
        if (Bk == 0)
            continue;
        const int64_t c0 = chgBase[b0], c1 = chgBase[b1];
        const size_t nc = (size_t)(c1 - c0);
        MTT_CUDA(ctx, cudaMemcpyAsync(dIH + b0 * nI, intHeight + b0 * nI, (size_t)Bk * nI * 8,
                                      cudaMemcpyHostToDevice, cs));
        MTT_CUDA(ctx, cudaMemcpyAsync(dIT + b0 * nI, intType + b0 * nI, (size_t)Bk * nI,
                                      cudaMemcpyHostToDevice, cs));
        if (nc > 0) {
            MTT_CUDA(ctx, cudaMemcpyAsync(dCT + c0, chgTime + c0, nc * 8, cudaMemcpyHostToDevice, cs));
            MTT_CUDA(ctx, cudaMemcpyAsync(dCTy + c0, chgType + c0, nc, cudaMemcpyHostToDevice, cs));
            MTT_CUDA(ctx, cudaMemcpyAsync(dO + c0, chgNode + c0, nc * 2, cudaMemcpyHostToDevice, cs));
        }
        if (parent)
            MTT_CUDA(ctx, cudaMemcpyAsync(dP + b0 * nN, parent + b0 * nN, (size_t)Bk * nN * 4,
                                          cudaMemcpyHostToDevice, cs));
        if (nChunks > 1) {
            MTT_CUDA(ctx, cudaEventRecord(ctx->events[k], cs));
            MTT_CUDA(ctx, cudaStreamWaitEvent(s, ctx->events[k], 0));
        }
        const long long total = (long long)Bk * nN;
        const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 32);
        k1_expand_tips<<<grid, 256, 0, s>>>(Bk, nL, dTipH, dTipT, dIH, dIT, b0, dH, dT,
                                            (long long *)dNodeOff);
        MTT_CHECK_LAUNCH(ctx);
        const bool perTree = nModels > 1;
        int rc = mtt_k1_launch(ctx, Bk, d, perTree ? Bk : 1, dM + (perTree ? b0 * d * d : 0),
                               dN + (perTree ? b0 * d : 0), dNodeOff + b0, dH, dP, dT, dO,
                               dChgBase + b0, dCT, dCTy, checkValidity, dLogP + b0,
                               dValid ? dValid + b0 : nullptr, maxEv);
        if (rc != MTT_OK)
            return rc;
        MTT_CUDA(ctx, cudaMemcpyAsync(logP + b0, dLogP + b0, Bk * 8, cudaMemcpyDeviceToHost, s));
        if (dValid)
            MTT_CUDA(ctx, cudaMemcpyAsync(valid + b0, dValid + b0, Bk, cudaMemcpyDeviceToHost, s));
    }
    MTT_CUDA(ctx, cudaStreamSynchronize(s));
    if (valid && !parent)
        memset(valid, 1, (size_t)B);
    return MTT_OK;
}

extern "C" int mtt_sc_set_max_tree_events(mtt_ctx *ctx, int64_t maxEvents)
{
    if (!ctx)
        return MTT_ERR_INVALID_ARG;
    ctx->k1MaxEvents = maxEvents;
    return MTT_OK;
}
