/*
 * MttNative.java -- Panama FFM (java.lang.foreign, final since JDK 22) binding of the C-ABI in
 * include/mtt_b200.h.  This is the ONLY place the drop-in plugins touch native code; no JNI
 * glue is needed.  NOT compiled in the build container (no JDK there): see INTEGRATION.md.
 *
 * One MttNative instance owns one mtt_ctx (one per chain / thread; a ctx is not re-entrant).
 */
package b200mtt;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.*;

public final class MttNative implements AutoCloseable {

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB =
            SymbolLookup.libraryLookup(System.getProperty("mtt.b200.lib", "libmtt_b200.so"), Arena.global());

    private static MethodHandle h(String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(
                () -> new UnsatisfiedLinkError("libmtt_b200.so does not export " + name)), fd);
    }

    private static final MethodHandle CTX_CREATE = h("mtt_ctx_create", FunctionDescriptor.of(JAVA_INT, JAVA_INT, ADDRESS));
    private static final MethodHandle CTX_DESTROY = h("mtt_ctx_destroy", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    private static final MethodHandle LAST_ERROR = h("mtt_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    private static final MethodHandle SC_LOGP = h("mtt_sc_logp_owner_batch", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, JAVA_LONG, JAVA_INT, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS,
            ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle UNIF_SET = h("mtt_unif_set_models", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_INT));
    private static final MethodHandle UNIF_BRANCH_LOGP = h("mtt_unif_branch_logp_batch", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle UNIF_RETYPE = h("mtt_unif_retype_batch", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS,
            JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle UNTYPED_LOGP = h("mtt_untyped_logp_batch", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, JAVA_LONG, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS,
            ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle UNTYPED_SAMPLE = h("mtt_untyped_sample_batch", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, JAVA_LONG, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS,
            ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle PEEL_CREATE = h("mtt_peel_create", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, JAVA_INT, JAVA_LONG, JAVA_INT, ADDRESS));
    private static final MethodHandle PEEL_DESTROY = h("mtt_peel_destroy", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    private static final MethodHandle PEEL_SET_STATES = h("mtt_peel_set_tip_states", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle PEEL_SET_WEIGHTS = h("mtt_peel_set_pattern_weights", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle PEEL_SET_SITE = h("mtt_peel_set_site_model", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_DOUBLE, JAVA_DOUBLE));
    private static final MethodHandle PEEL_SET_HKY = h("mtt_peel_set_hky", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_DOUBLE));
    private static final MethodHandle PEEL_LOGL = h("mtt_peel_logl_batch", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle CHAINS_CREATE = h("mtt_peel_chains_create", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, JAVA_INT, ADDRESS));
    private static final MethodHandle CHAINS_DESTROY = h("mtt_peel_chains_destroy", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    private static final MethodHandle CHAINS_UPDATE = h("mtt_peel_chains_update", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle CHAINS_UPDATE_SPARSE = h("mtt_peel_chains_update_sparse", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle CHAINS_ACCEPT = h("mtt_peel_chains_accept", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, ADDRESS));

    /** status codes of mtt_unif_* (MTT_ST_*) */
    public static final int ST_OK = 0, ST_NO_VALID_PATH = 1, ST_PBA_RUBBISH = 2, ST_FB_FELL_THROUGH = 3,
            ST_NEEDS_POISSON_TAIL = 4, ST_NEED_MORE_UNIFORMS = 5, ST_SKIPPED = 7;

    private final MemorySegment ctx;

    public MttNative(int device) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(ADDRESS);
            int rc = (int) CTX_CREATE.invoke(device, out);
            if (rc != 0)
                throw new IllegalStateException("mtt_ctx_create failed (" + rc + "): " + lastError(MemorySegment.NULL));
            ctx = out.get(ADDRESS, 0);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    static String lastError(MemorySegment ctx) {
        try {
            MemorySegment s = (MemorySegment) LAST_ERROR.invoke(ctx);
            return s.reinterpret(512).getString(0);
        } catch (Throwable t) {
            return "<no message>";
        }
    }

    void check(int rc) {
        if (rc != 0)
            throw new IllegalStateException("libmtt_b200 error " + rc + ": " + lastError(ctx));
    }

    /** mtt_sc_logp_owner_batch: all segments are off-heap (Arena) buffers in the layout of mtt_b200.h. */
    public void scLogPBatch(long B, int d, long nModels, MemorySegment m, MemorySegment N,
                            MemorySegment nodeOff, MemorySegment height, MemorySegment parent,
                            MemorySegment nodeType, MemorySegment chgNode, MemorySegment chgBase,
                            MemorySegment chgTime, MemorySegment chgType, boolean checkValidity,
                            MemorySegment logP, MemorySegment valid) {
        try {
            check((int) SC_LOGP.invoke(ctx, B, d, nModels, m, N, nodeOff, height, parent, nodeType, chgNode,
                    chgBase, chgTime, chgType, checkValidity ? 1 : 0, logP, valid));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    public void unifSetModels(int nModels, int d, MemorySegment m, boolean sym) {
        try {
            check((int) UNIF_SET.invoke(ctx, nModels, d, m, sym ? 1 : 0));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    public void unifBranchLogP(long n, MemorySegment model, MemorySegment tBottom, MemorySegment tTop,
                               MemorySegment typeBottom, MemorySegment typeTop, MemorySegment chgOff,
                               MemorySegment chgTime, MemorySegment chgType, MemorySegment logProb,
                               MemorySegment status) {
        try {
            check((int) UNIF_BRANCH_LOGP.invoke(ctx, n, model, tBottom, tTop, typeBottom, typeTop, chgOff, chgTime,
                    chgType, logProb, status));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    public void unifRetype(long nJobs, MemorySegment brOff, MemorySegment model, MemorySegment tBottom,
                           MemorySegment tTop, MemorySegment typeBottom, MemorySegment typeTop,
                           MemorySegment nVirtForced, MemorySegment uniforms, MemorySegment uOff, int outCap,
                           MemorySegment nChg, MemorySegment outChgTime, MemorySegment outChgType,
                           MemorySegment logProb, MemorySegment status, MemorySegment nVirt, MemorySegment uUsed) {
        try {
            check((int) UNIF_RETYPE.invoke(ctx, nJobs, brOff, model, tBottom, tTop, typeBottom, typeTop, nVirtForced,
                    uniforms, uOff, outCap, nChg, outChgTime, outChgType, logProb, status, nVirt, uUsed));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    /** mtt_untyped_logp_batch: nTrees x nParticles importance-sampling particles, one call. */
    public void untypedLogPBatch(long nTrees, int nParticles, MemorySegment model, MemorySegment N,
                                 MemorySegment nodeOff, MemorySegment height, MemorySegment parent,
                                 MemorySegment leafType, MemorySegment leafOrder, MemorySegment uniforms,
                                 MemorySegment uOff, MemorySegment logP, MemorySegment logW,
                                 MemorySegment status, MemorySegment uUsed) {
        try {
            check((int) UNTYPED_LOGP.invoke(ctx, nTrees, nParticles, model, N, nodeOff, height, parent, leafType,
                    leafOrder, uniforms, uOff, logP, logW, status, uUsed));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    /** mtt_untyped_sample_batch: the particles returned as typings (node types + change lists). */
    public void untypedSampleBatch(long nTrees, int nParticles, MemorySegment model, MemorySegment N,
                                   MemorySegment nodeOff, MemorySegment height, MemorySegment parent,
                                   MemorySegment leafType, MemorySegment leafOrder, MemorySegment uniforms,
                                   MemorySegment uOff, int chgCap, MemorySegment nodeType, MemorySegment nChg,
                                   MemorySegment chgTime, MemorySegment chgType, MemorySegment chgNode,
                                   MemorySegment logW, MemorySegment status, MemorySegment uUsed) {
        try {
            check((int) UNTYPED_SAMPLE.invoke(ctx, nTrees, nParticles, model, N, nodeOff, height, parent, leafType,
                    leafOrder, uniforms, uOff, chgCap, nodeType, nChg, chgTime, chgType, chgNode, logW, status,
                    uUsed));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    /** Peeling handle (mtt_peel_*). */
    public final class Peel implements AutoCloseable {
        private final MemorySegment h;

        Peel(int nTips, long nPatterns, int nCat) {
            try (Arena a = Arena.ofConfined()) {
                MemorySegment out = a.allocate(ADDRESS);
                check((int) PEEL_CREATE.invoke(ctx, nTips, nPatterns, nCat, out));
                h = out.get(ADDRESS, 0);
            } catch (RuntimeException e) {
                throw e;
            } catch (Throwable t) {
                throw new RuntimeException(t);
            }
        }

        public void setTipStates(MemorySegment states) { call(PEEL_SET_STATES, h, states); }
        public void setPatternWeights(MemorySegment w) { call(PEEL_SET_WEIGHTS, h, w); }
        public void setSiteModel(MemorySegment rates, MemorySegment props, MemorySegment freqs, double pInv, double clock) {
            call(PEEL_SET_SITE, h, rates, props, freqs, pInv, clock);
        }
        public void setHky(double kappa) { call(PEEL_SET_HKY, h, kappa); }
        public void logLBatch(long B, MemorySegment parent, MemorySegment height, MemorySegment logL) {
            call(PEEL_LOGL, h, B, parent, height, logL);
        }
        /** Resident chains on this handle's alignment (mtt_peel_chains_*). */
        public Chains newChains(int nChains) { return new Chains(h, nChains); }
        @Override public void close() { call(PEEL_DESTROY, h); }
    }

    /** Partials resident in HBM, dirty-node updates, accept / reject (mtt_peel_chains_*). */
    public final class Chains implements AutoCloseable {
        private final MemorySegment s;

        Chains(MemorySegment peel, int nChains) {
            try (Arena a = Arena.ofConfined()) {
                MemorySegment out = a.allocate(ADDRESS);
                check((int) CHAINS_CREATE.invoke(peel, nChains, out));
                s = out.get(ADDRESS, 0);
            } catch (RuntimeException e) {
                throw e;
            } catch (Throwable t) {
                throw new RuntimeException(t);
            }
        }

        /** branchRates, kappa, clockRate, dirty may be MemorySegment.NULL */
        public void update(MemorySegment parent, MemorySegment height, MemorySegment branchRates,
                           MemorySegment kappa, MemorySegment clockRate, MemorySegment dirty, MemorySegment logL) {
            call(CHAINS_UPDATE, s, parent, height, branchRates, kappa, clockRate, dirty, logL);
        }
        /** the proposal as lists of changed nodes: off[nChains+1], then node / parent / height per entry */
        public void updateSparse(MemorySegment off, MemorySegment node, MemorySegment parent, MemorySegment height,
                                 MemorySegment branchRates, MemorySegment kappa, MemorySegment clockRate,
                                 MemorySegment logL) {
            call(CHAINS_UPDATE_SPARSE, s, off, node, parent, height, branchRates, kappa, clockRate, logL);
        }
        public void accept(MemorySegment mask) { call(CHAINS_ACCEPT, s, mask); }
        @Override public void close() { call(CHAINS_DESTROY, s); }
    }

    public Peel newPeel(int nTips, long nPatterns, int nCat) { return new Peel(nTips, nPatterns, nCat); }

    private void call(MethodHandle mh, Object... args) {
        try {
            check((int) mh.invokeWithArguments(args));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    @Override
    public void close() {
        try {
            CTX_DESTROY.invoke(ctx);
        } catch (Throwable ignored) {
        }
    }
}
