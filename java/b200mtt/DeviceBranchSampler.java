/*
 * DeviceBranchSampler -- the two protected methods of
 * multitypetree.operators.UniformizationRetypeOperator on the device, one branch per call:
 *   retypeBranch(srcNode)       (UniformizationRetypeOperator.java:116-225) -> mtt_unif_retype_batch
 *   getBranchTypeProb(srcNode)  (:233-275)                                  -> mtt_unif_branch_logp_batch
 * Every operator that extends UniformizationRetypeOperator reaches the uniformization code only
 * through these two methods (NodeShiftRetype.java:83-137, TypedSubtreeExchange.java:75-84,
 * TypedWilsonBalding.java:87-289), so the shims b200mtt.NodeShiftRetype / TypedSubtreeExchange /
 * TypedWilsonBalding subclass the reference's operators, keep their tree surgery untouched, and
 * override just these two.  Randomizer order is the reference's: the uniforms are drawn from
 * BEAST's Randomizer as a window, the kernel reports how many it consumed, the stream is put back
 * to exactly that position (RandomizerState); the steady-state Poisson tail (:96-102) is drawn
 * here with Randomizer.nextPoisson and the branch resubmitted with the forced event count.
 * One branch per device call is a latency path (about 25 us); batches of chains go through
 * mtt_unif_node_retype_batch (include/mtt_b200.hpp, NodeRetypeEnsemble).
 * NOT compiled here (no JDK in the build container).
 */
package b200mtt;

import beast.base.util.Randomizer;
import multitypetree.evolution.tree.MigrationModel;
import multitypetree.evolution.tree.MultiTypeNode;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;

import static java.lang.foreign.ValueLayout.*;

final class DeviceBranchSampler implements AutoCloseable {

    /** retype(): the branch has no valid path (Pba == 0): the caller throws NoValidPathException. */
    static final double NO_VALID_PATH = Double.NaN;

    private MttNative nativeLib;
    private Arena arena;
    private MigrationModel migModel;
    private boolean sym;
    private int outCap = 256;
    private MemorySegment mDense, brOff, model, t0, t1, a, b, forced, uOff, chgOff, nChg, logProb, status,
            nVirt, uUsed, chgTime, chgType, outT, outTy;

    void init(MigrationModel migModel, boolean useSymmetrizedRates) {
        this.migModel = migModel;
        this.sym = useSymmetrizedRates;
        nativeLib = new MttNative(-1);
        arena = Arena.ofShared();
        int d = migModel.getNTypes();
        mDense = arena.allocate(8L * d * d, 16);
        brOff = arena.allocate(16, 8);
        model = arena.allocate(4, 4);
        t0 = arena.allocate(8, 8);
        t1 = arena.allocate(8, 8);
        a = arena.allocate(1, 1);
        b = arena.allocate(1, 1);
        forced = arena.allocate(4, 4);
        uOff = arena.allocate(16, 8);
        chgOff = arena.allocate(16, 8);
        nChg = arena.allocate(4, 4);
        logProb = arena.allocate(8, 8);
        status = arena.allocate(4, 4);
        nVirt = arena.allocate(4, 4);
        uUsed = arena.allocate(8, 8);
        allocOut();
    }

    private void allocOut() {
        outT = arena.allocate(8L * outCap, 8);
        outTy = arena.allocate(outCap, 1);
    }

    /** The migration model is a CalculationNode: its parameters may have changed since the last proposal. */
    private void pushModel() {
        int d = migModel.getNTypes();
        for (int i = 0; i < d; i++)
            for (int j = 0; j < d; j++)
                mDense.setAtIndex(JAVA_DOUBLE, (long) i * d + j, i == j ? 0.0 : migModel.getBackwardRate(i, j));
        nativeLib.unifSetModels(1, d, mDense, sym);
    }

    private void describe(MultiTypeNode s) {
        model.set(JAVA_INT, 0, 0);
        t0.set(JAVA_DOUBLE, 0, s.getHeight());
        t1.set(JAVA_DOUBLE, 0, s.getParent().getHeight());
        a.set(JAVA_BYTE, 0, (byte) s.getNodeType());
        b.set(JAVA_BYTE, 0, (byte) ((MultiTypeNode) s.getParent()).getNodeType());
    }

    /** getBranchTypeProb (:233-275): log probability of the branch's current typing; no random numbers. */
    double branchTypeProb(MultiTypeNode s) {
        pushModel();
        describe(s);
        int c = s.getChangeCount();
        if (chgTime == null || chgTime.byteSize() < 8L * Math.max(c, 1)) {
            chgTime = arena.allocate(8L * Math.max(2 * c, 64), 8);
            chgType = arena.allocate(Math.max(2 * c, 64), 1);
        }
        for (int k = 0; k < c; k++) {
            chgTime.setAtIndex(JAVA_DOUBLE, k, s.getChangeTime(k));
            chgType.set(JAVA_BYTE, k, (byte) s.getChangeType(k));
        }
        chgOff.setAtIndex(JAVA_LONG, 0, 0);
        chgOff.setAtIndex(JAVA_LONG, 1, c);
        nativeLib.unifBranchLogP(1, model, t0, t1, a, b, chgOff, chgTime, chgType, logProb, status);
        // MTT_ST_PBA_RUBBISH: the reference warns and returns -Infinity (:267-270)
        return status.get(JAVA_INT, 0) == MttNative.ST_OK ? logProb.get(JAVA_DOUBLE, 0) : Double.NEGATIVE_INFINITY;
    }

    /**
     * retypeBranch (:116-225): draws a new typing of the branch above s conditional on the types at its
     * two ends, installs it on s, returns its log probability; NO_VALID_PATH (NaN) when Pba == 0 (:137-138),
     * -Infinity on the reference's numerical guards (:141-144, :179-190), the branch then untouched.
     */
    double retype(MultiTypeNode s) {
        pushModel();
        describe(s);
        brOff.setAtIndex(JAVA_LONG, 0, 0);
        brOff.setAtIndex(JAVA_LONG, 1, 1);
        int nForced = -1;
        long window = 256;
        for (;;) {
            forced.set(JAVA_INT, 0, nForced);
            // BEAST's Randomizer has no peek: draw the window, put the stream back, skip what was used
            long[] saved = RandomizerState.save();
            MemorySegment uni = arena.allocate(8L * window, 8);
            for (long i = 0; i < window; i++)
                uni.setAtIndex(JAVA_DOUBLE, i, Randomizer.nextDouble());
            uOff.setAtIndex(JAVA_LONG, 0, 0);
            uOff.setAtIndex(JAVA_LONG, 1, window);
            nativeLib.unifRetype(1, brOff, model, t0, t1, a, b, forced, uni, uOff, outCap, nChg, outT, outTy,
                    logProb, status, nVirt, uUsed);
            int st = status.get(JAVA_INT, 0);
            if (st == MttNative.ST_NEED_MORE_UNIFORMS) {           // nothing consumed: same draw, longer window
                RandomizerState.restoreAndSkip(saved, 0);
                window *= 4;
                continue;
            }
            RandomizerState.restoreAndSkip(saved, uUsed.get(JAVA_LONG, 0));
            if (st == MttNative.ST_NEEDS_POISSON_TAIL) {           // :96-102, thresh = the count reached
                int thresh = nVirt.get(JAVA_INT, 0);
                double muL = migModel.getMu(sym) * (t1.get(JAVA_DOUBLE, 0) - t0.get(JAVA_DOUBLE, 0));
                do {
                    nForced = (int) Randomizer.nextPoisson(muL);
                } while (nForced < thresh);
                continue;
            }
            if (st == MttNative.ST_NO_VALID_PATH)
                return NO_VALID_PATH;
            if (st != MttNative.ST_OK)
                return Double.NEGATIVE_INFINITY;
            int c = nChg.get(JAVA_INT, 0);
            if (c > outCap) {                                       // more changes than room: replay the draw
                outCap = 2 * c;
                allocOut();
                RandomizerState.restoreAndSkip(saved, 0);
                continue;
            }
            s.clearChanges();
            for (int k = 0; k < c; k++)
                s.addChange(outTy.get(JAVA_BYTE, k), outT.getAtIndex(JAVA_DOUBLE, k));
            return logProb.get(JAVA_DOUBLE, 0);
        }
    }

    @Override
    public void close() {
        if (arena != null)
            arena.close();
        if (nativeLib != null)
            nativeLib.close();
    }
}
