/*
 * Drop-in replacement of beast.base.evolution.likelihood.TreeLikelihood for the configuration the
 * reference wires (examples/structuredCoalescent.xml:568-572; fxtemplates/MultiTypeTree.xml:184-197):
 * Inputs "data", "tree", "siteModel", "branchRateModel" are inherited from GenericTreeLikelihood;
 * nucleotide data, HKY, gamma categories / proportion invariant, strict clock.  Peeling runs on the
 * B200 on ONE resident chain (mtt_peel_chains_*): the partials of every internal node stay in HBM,
 * calculateLogP() recomputes the nodes above the dirty ones (Node.isDirty(), as
 * TreeLikelihood.traverse [EXT]), accept() / restore() keep or drop the proposal -- the roles of
 * BeerLikelihoodCore's currentPartialsIndex / storedPartialsIndex.  NOT compiled here (no JDK).
 */
package b200mtt;

import beast.base.core.Description;
import beast.base.evolution.alignment.Alignment;
import beast.base.evolution.branchratemodel.BranchRateModel;
import beast.base.evolution.likelihood.GenericTreeLikelihood;
import beast.base.evolution.sitemodel.SiteModel;
import beast.base.evolution.substitutionmodel.HKY;
import beast.base.evolution.tree.Node;
import beast.base.evolution.tree.TreeInterface;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;

import static java.lang.foreign.ValueLayout.*;

@Description("Felsenstein pruning on a B200 for 4-state data under HKY (MultiTypeTree hot path).")
public class TreeLikelihood extends GenericTreeLikelihood {

    private MttNative nativeLib;
    private MttNative.Peel peel;
    private MttNative.Chains chain;
    private Arena arena;
    private MemorySegment parent, height, dirtyNode, dirtyParent, dirtyHeight, off, one, zero, rates, props, freqs, out;
    private boolean needFull = true, pushedThisProposal = false;
    private double[] lastRates, lastProps, lastFreqs;
    private double lastPInv = Double.NaN, lastClock = Double.NaN, lastKappa = Double.NaN;
    private int nTips, nCat;

    @Override
    public void initAndValidate() {
        Alignment data = dataInput.get();
        TreeInterface tree = treeInput.get();
        SiteModel site = (SiteModel) siteModelInput.get();
        if (!(site.substModelInput.get() instanceof HKY) || data.getMaxStateCount() != 4)
            throw new IllegalArgumentException("b200mtt.TreeLikelihood supports nucleotide data under HKY");
        site.setDataType(data.getDataType());
        site.setPropInvariantIsCategory(false);       // as TreeLikelihood does [EXT]
        nTips = tree.getLeafNodeCount();
        nCat = site.getCategoryCount();
        int P = data.getPatternCount();
        nativeLib = new MttNative(-1);
        peel = nativeLib.newPeel(nTips, P, nCat);
        arena = Arena.ofShared();
        MemorySegment states = arena.allocate(4L * nTips * P, 16);
        MemorySegment weights = arena.allocate(4L * P, 16);
        for (int t = 0; t < nTips; t++) {
            int taxon = data.getTaxonIndex(tree.getNode(t).getID());
            for (int p = 0; p < P; p++) {
                int s = data.getPattern(taxon, p);
                states.setAtIndex(JAVA_INT, (long) t * P + p, data.getDataType().isAmbiguousCode(s) ? 4 : s);
            }
        }
        for (int p = 0; p < P; p++)
            weights.setAtIndex(JAVA_INT, p, data.getPatternWeight(p));
        peel.setTipStates(states);
        peel.setPatternWeights(weights);
        int N = 2 * nTips - 1;
        parent = arena.allocate(4L * N, 16); height = arena.allocate(8L * N, 16);
        rates = arena.allocate(8L * nCat, 8); props = arena.allocate(8L * nCat, 8);
        freqs = arena.allocate(32, 8); out = arena.allocate(8, 8);
        dirtyNode = arena.allocate(4L * N, 16); dirtyParent = arena.allocate(4L * N, 16);
        dirtyHeight = arena.allocate(8L * N, 16); off = arena.allocate(16, 8);
        one = arena.allocate(1, 1); zero = arena.allocate(1, 1);
        one.set(JAVA_BYTE, 0, (byte) 1);
        chain = peel.newChains(1);
    }

    @Override
    public double calculateLogP() {
        SiteModel site = (SiteModel) siteModelInput.get();
        HKY hky = (HKY) site.substModelInput.get();
        BranchRateModel clock = branchRateModelInput.get();
        TreeInterface tree = treeInput.get();
        double[] r = site.getCategoryRates(null), pr = site.getCategoryProportions(null), f = hky.getFrequencies();
        for (int c = 0; c < nCat; c++) { rates.setAtIndex(JAVA_DOUBLE, c, r[c]); props.setAtIndex(JAVA_DOUBLE, c, pr[c]); }
        for (int i = 0; i < 4; i++) freqs.setAtIndex(JAVA_DOUBLE, i, f[i]);
        double clockRate = clock == null ? 1.0 : clock.getRateForBranch(tree.getRoot().getLeft());   // strict clock
        double kappa = hky.kappaInput.get().getValue(), pInv = site.getProportionInvariant();
        // the setters make every resident partial stale: only call them when a value moved
        if (!java.util.Arrays.equals(r, lastRates) || !java.util.Arrays.equals(pr, lastProps)
                || !java.util.Arrays.equals(f, lastFreqs) || pInv != lastPInv || clockRate != lastClock) {
            peel.setSiteModel(rates, props, freqs, pInv, clockRate);
            lastRates = r.clone(); lastProps = pr.clone(); lastFreqs = f.clone(); lastPInv = pInv; lastClock = clockRate;
            needFull = true; pushedThisProposal = true;
        }
        if (kappa != lastKappa) { peel.setHky(kappa); lastKappa = kappa; needFull = true; pushedThisProposal = true; }
        if (needFull) {
            // first evaluation, or a model parameter moved: the whole tree, every node recomputed
            for (Node node : tree.getNodesAsArray()) {
                int nr = node.getNr();
                parent.setAtIndex(JAVA_INT, nr, node.isRoot() ? -1 : node.getParent().getNr());
                height.setAtIndex(JAVA_DOUBLE, nr, node.getHeight());
            }
            chain.update(parent, height, MemorySegment.NULL, MemorySegment.NULL, MemorySegment.NULL,
                    MemorySegment.NULL, out);
            needFull = false;
        } else {
            // a tree move: only the dirty nodes cross the boundary; they are compared with the
            // accepted tree and the nodes above a real change are recomputed
            long c = 0;
            for (Node node : tree.getNodesAsArray()) {
                if (node.isDirty() == beast.base.evolution.tree.Tree.IS_CLEAN)
                    continue;
                dirtyNode.setAtIndex(JAVA_INT, c, node.getNr());
                dirtyParent.setAtIndex(JAVA_INT, c, node.isRoot() ? -1 : node.getParent().getNr());
                dirtyHeight.setAtIndex(JAVA_DOUBLE, c, node.getHeight());
                c++;
            }
            off.setAtIndex(JAVA_LONG, 0, 0); off.setAtIndex(JAVA_LONG, 1, c);
            chain.updateSparse(off, dirtyNode, dirtyParent, dirtyHeight, MemorySegment.NULL, MemorySegment.NULL,
                    MemorySegment.NULL, out);
        }
        logP = out.get(JAVA_DOUBLE, 0);
        return logP;
    }

    /** BEAST calls store() before a proposal: the library's accepted state already is the stored one. */
    @Override
    public void store() { super.store(); }

    /** proposal rejected: drop it (BeerLikelihoodCore.restore).  A rejected model-parameter move
     *  leaves the library's setters at the rejected value: forget what was pushed, so that the
     *  next evaluation pushes the restored values (and is a full one); a rejected tree move
     *  costs nothing */
    @Override
    public void restore() {
        chain.accept(zero);
        if (pushedThisProposal) { lastKappa = Double.NaN; lastPInv = Double.NaN; }
        pushedThisProposal = false;
        super.restore();
    }

    /** proposal accepted: its partials become the accepted ones */
    @Override
    protected void accept() { chain.accept(one); pushedThisProposal = false; super.accept(); }

    @Override
    protected boolean requiresRecalculation() {
        return true;      // the library decides what to recompute (nothing when nothing changed)
    }
}
