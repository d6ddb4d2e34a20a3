/*
 * Drop-in replacement of multitypetree.operators.NodeRetype (operators/NodeRetype.java:28-65)
 * on top of the uniformization entry points: same Inputs ("multiTypeTree", "migrationModel" from
 * MultiTypeTreeOperator.java:37-45, "useSymmetrizedRates" from
 * UniformizationRetypeOperator.java:41-43, "weight"), same Randomizer call order
 * (nextInt(internal), nextInt(nTypes), then the doubles of retypeBranch in branch order
 * node / left / right).  The uniforms are drawn from BEAST's Randomizer on the Java side and
 * handed to the kernel as a window; the kernel reports how many it consumed.
 * NOT compiled here (no JDK).  Sketch-level completeness: the Poisson tail and the
 * NEED_MORE_UNIFORMS retry follow multitypetree_b200/uniformization.py::NodeRetype.proposal.
 */
package b200mtt;

import beast.base.core.Description;
import beast.base.evolution.tree.Node;
import beast.base.util.Randomizer;
import multitypetree.evolution.tree.MultiTypeNode;
import multitypetree.operators.UniformizationRetypeOperator;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;

import static java.lang.foreign.ValueLayout.*;

@Description("Retypes a randomly chosen node and its attached branches (B200-native uniformization).")
public class NodeRetype extends UniformizationRetypeOperator {

    private static final int WINDOW = 256, OUT_CAP = 256;
    private MttNative nativeLib;
    private Arena arena;
    private MemorySegment mDense, brOff, model, t0, t1, a, b, forced, uni, uOff, chgOff, chgTime, chgType,
            nChg, outT, outTy, logProb, status, nVirt, uUsed;
    private boolean modelDirty = true;

    @Override
    public void initAndValidate() {
        super.initAndValidate();
        nativeLib = new MttNative(-1);
        arena = Arena.ofShared();
        int d = migModel.getNTypes();
        mDense = arena.allocate(8L * d * d, 16);
        brOff = arena.allocate(16, 8); model = arena.allocate(12, 4);
        t0 = arena.allocate(24, 8); t1 = arena.allocate(24, 8);
        a = arena.allocate(3, 1); b = arena.allocate(3, 1); forced = arena.allocate(12, 4);
        uni = arena.allocate(8L * WINDOW * 16, 8); uOff = arena.allocate(16, 8);
        chgOff = arena.allocate(32, 8); chgTime = arena.allocate(8L * 4096, 8); chgType = arena.allocate(4096, 1);
        nChg = arena.allocate(12, 4); outT = arena.allocate(8L * 3 * OUT_CAP, 8); outTy = arena.allocate(3 * OUT_CAP, 1);
        logProb = arena.allocate(24, 8); status = arena.allocate(12, 4); nVirt = arena.allocate(12, 4);
        uUsed = arena.allocate(8, 8);
    }

    private void pushModel() {
        int d = migModel.getNTypes();
        for (int i = 0; i < d; i++)
            for (int j = 0; j < d; j++)
                mDense.setAtIndex(JAVA_DOUBLE, (long) i * d + j, i == j ? 0.0 : migModel.getBackwardRate(i, j));
        nativeLib.unifSetModels(1, d, mDense, useSymmetrizedRatesInput.get());   // drops cached R^k
    }

    @Override
    public double proposal() {
        // the migration model is a CalculationNode: re-push when its parameters changed
        pushModel();
        double logHR = 0.0;
        Node node = mtTree.getNode(mtTree.getLeafNodeCount() + Randomizer.nextInt(mtTree.getInternalNodeCount()));
        Node[] br = node.isRoot() ? new Node[]{node.getLeft(), node.getRight()}
                                  : new Node[]{node, node.getLeft(), node.getRight()};
        int nb = br.length;

        // old path probabilities (getBranchTypeProb, :233-275).  (mtt_unif_node_retype_batch does
        // this and the retyping below in ONE device call -- it needs a second set of branch
        // buffers for the old node types; see mtt::NodeRetypeEnsemble in include/mtt_b200.hpp.)
        long c = 0;
        for (int i = 0; i < nb; i++) {
            MultiTypeNode s = (MultiTypeNode) br[i];
            describe(i, s);
            chgOff.setAtIndex(JAVA_LONG, i, c);
            for (int k = 0; k < s.getChangeCount(); k++, c++) {
                chgTime.setAtIndex(JAVA_DOUBLE, c, s.getChangeTime(k));
                chgType.set(JAVA_BYTE, c, (byte) s.getChangeType(k));
            }
        }
        chgOff.setAtIndex(JAVA_LONG, nb, c);
        nativeLib.unifBranchLogP(nb, model, t0, t1, a, b, chgOff, chgTime, chgType, logProb, status);
        for (int i = 0; i < nb; i++)
            logHR += logProb.getAtIndex(JAVA_DOUBLE, i);

        ((MultiTypeNode) node).setNodeType(Randomizer.nextInt(migModel.getNTypes()));

        // retype (retypeBranch, :116-225): one job, one uniform window
        for (int i = 0; i < nb; i++) {
            describe(i, (MultiTypeNode) br[i]);
            forced.setAtIndex(JAVA_INT, i, -1);
        }
        brOff.setAtIndex(JAVA_LONG, 0, 0); brOff.setAtIndex(JAVA_LONG, 1, nb);
        // BEAST's Randomizer has no peek: draw the window, and push back what was not consumed by
        // re-seeding from a saved state (Randomizer.getState/setState [EXT]) -- see INTEGRATION.md.
        long[] saved = RandomizerState.save();
        int window = WINDOW;
        for (int i = 0; i < window; i++)
            uni.setAtIndex(JAVA_DOUBLE, i, Randomizer.nextDouble());
        uOff.setAtIndex(JAVA_LONG, 0, 0); uOff.setAtIndex(JAVA_LONG, 1, window);
        nativeLib.unifRetype(1, brOff, model, t0, t1, a, b, forced, uni, uOff, OUT_CAP, nChg, outT, outTy,
                logProb, status, nVirt, uUsed);
        RandomizerState.restoreAndSkip(saved, uUsed.get(JAVA_LONG, 0));

        double newLp = 0.0;
        for (int i = 0; i < nb; i++) {
            int st = status.getAtIndex(JAVA_INT, i);
            if (st == MttNative.ST_NO_VALID_PATH)
                return Double.NEGATIVE_INFINITY;                    // NoValidPathException (:57-58)
            if (st != MttNative.ST_OK)
                throw new IllegalStateException("retype status " + st + ": resubmit (see uniformization.py)");
            MultiTypeNode s = (MultiTypeNode) br[i];
            s.clearChanges();
            for (int k = 0; k < nChg.getAtIndex(JAVA_INT, i); k++)
                s.addChange(outTy.get(JAVA_BYTE, (long) i * OUT_CAP + k),
                        outT.getAtIndex(JAVA_DOUBLE, (long) i * OUT_CAP + k));
            newLp += logProb.getAtIndex(JAVA_DOUBLE, i);
        }
        return logHR - newLp;
    }

    private void describe(int i, MultiTypeNode s) {
        model.setAtIndex(JAVA_INT, i, 0);
        t0.setAtIndex(JAVA_DOUBLE, i, s.getHeight());
        t1.setAtIndex(JAVA_DOUBLE, i, s.getParent().getHeight());
        a.set(JAVA_BYTE, i, (byte) s.getNodeType());
        b.set(JAVA_BYTE, i, (byte) ((MultiTypeNode) s.getParent()).getNodeType());
    }
}
