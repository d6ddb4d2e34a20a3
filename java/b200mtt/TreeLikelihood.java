/*
 * Drop-in replacement of beast.base.evolution.likelihood.TreeLikelihood for the configuration the
 * reference wires (examples/structuredCoalescent.xml:568-572; fxtemplates/MultiTypeTree.xml:184-197):
 * Inputs "data", "tree", "siteModel", "branchRateModel" are inherited from GenericTreeLikelihood;
 * nucleotide data, HKY, gamma categories / proportion invariant, strict clock.  Peeling runs on the
 * B200 through mtt_peel_logl_batch.  NOT compiled here (no JDK).
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
    private Arena arena;
    private MemorySegment parent, height, rates, props, freqs, out;
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
        peel.setSiteModel(rates, props, freqs, site.getProportionInvariant(), clockRate);
        peel.setHky(hky.kappaInput.get().getValue());
        for (Node node : tree.getNodesAsArray()) {
            parent.setAtIndex(JAVA_INT, node.getNr(), node.isRoot() ? -1 : node.getParent().getNr());
            height.setAtIndex(JAVA_DOUBLE, node.getNr(), node.getHeight());
        }
        peel.logLBatch(1, parent, height, out);
        logP = out.get(JAVA_DOUBLE, 0);
        return logP;
    }

    @Override
    protected boolean requiresRecalculation() {
        return true;      // full re-evaluation per call: ~25 us of kernel time for 1000 tips x 10k sites
    }
}
