/*
 * Drop-in replacement of multitypetree.distributions.StructuredCoalescentUntypedTreeDensity
 * (distributions/StructuredCoalescentUntypedTreeDensity.java:20-535): same Inputs ("tree" from
 * TreeDistribution, "migrationModel", "nParticles", "typeLabel"), same contract of
 * calculateLogP() (stochastic: every call draws new particles).  All nParticles particles run
 * in ONE call of mtt_untyped_logp_batch (one warp per particle).  The leaf orders
 * (Randomizer.shuffled, :129) and one window of uniforms per particle are drawn from BEAST's
 * Randomizer here, so a chain stays reproducible from its seed; windows are consumed whole, so the
 * stream position after a call differs from the particle-by-particle Java loop.
 * NOT compiled here (no JDK in the build container).
 */
package b200mtt;

import beast.base.core.Description;
import beast.base.core.Input;
import beast.base.core.Input.Validate;
import beast.base.evolution.tree.Node;
import beast.base.evolution.tree.TraitSet;
import beast.base.evolution.tree.Tree;
import beast.base.evolution.tree.TreeDistribution;
import beast.base.util.Randomizer;
import multitypetree.evolution.tree.SCMigrationModel;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.util.ArrayList;
import java.util.List;
import java.util.TreeSet;

import static java.lang.foreign.ValueLayout.*;

@Description("Structured-coalescent density of an untyped tree by particle importance sampling (B200-native).")
public class StructuredCoalescentUntypedTreeDensity extends TreeDistribution {

    public Input<SCMigrationModel> migrationModelInput = new Input<>(
            "migrationModel", "Model of migration between demes.", Validate.REQUIRED);
    public Input<Integer> nParticlesInput = new Input<>(
            "nParticles", "Number of simulated particles used to estimate the tree density.", Validate.REQUIRED);
    public Input<String> typeLabelInput = new Input<>(
            "typeLabel", "Label for type traits (default 'type')", "type");
    public Input<Integer> deviceInput = new Input<>("device", "CUDA device ordinal (default: current).", -1);

    private MttNative nativeLib;
    private Tree tree;
    private SCMigrationModel migrationModel;
    private int nParticles;
    private int[] leafTypes;

    @Override
    public void initAndValidate() {
        super.initAndValidate();
        nParticles = nParticlesInput.get();
        tree = (Tree) treeInput.get();
        migrationModel = migrationModelInput.get();
        // Leaf types: index of each leaf's trait value in the sorted set of distinct values of
        // the trait named typeLabel (same numbering as the reference, :88-111).
        final String label = typeLabelInput.get();
        TraitSet types = tree.m_traitList.get().stream()
                .filter(t -> t.getTraitName().equals(label)).findFirst()
                .orElseThrow(() -> new IllegalArgumentException(
                        "Trait set (with name '" + label + "') must be provided."));
        final TraitSet traits = types;
        List<String> sortedValues = new ArrayList<>(new TreeSet<>(
                java.util.Arrays.stream(tree.getTaxaNames()).map(traits::getStringValue).toList()));
        leafTypes = new int[tree.getLeafNodeCount()];
        for (int nr = 0; nr < leafTypes.length; nr++)
            leafTypes[nr] = sortedValues.indexOf(traits.getStringValue(tree.getNode(nr).getID()));
        nativeLib = new MttNative(deviceInput.get());
    }

    @Override
    public double calculateLogP() {
        final int d = migrationModel.getNTypes(), n = tree.getNodeCount(), nLeaves = tree.getLeafNodeCount();
        try (Arena a = Arena.ofConfined()) {
            MemorySegment m = a.allocate(8L * d * d, 16), N = a.allocate(8L * d, 16);
            for (int i = 0; i < d; i++) {
                N.setAtIndex(JAVA_DOUBLE, i, migrationModel.getPopSize(i));
                for (int j = 0; j < d; j++)
                    m.setAtIndex(JAVA_DOUBLE, (long) i * d + j, i == j ? 0.0 : migrationModel.getBackwardRate(i, j));
            }
            nativeLib.unifSetModels(1, d, m, false);
            MemorySegment nodeOff = a.allocate(16, 8), height = a.allocate(8L * n, 16),
                    parent = a.allocate(4L * n, 16), leafType = a.allocate(n, 16);
            nodeOff.setAtIndex(JAVA_LONG, 0, 0);
            nodeOff.setAtIndex(JAVA_LONG, 1, n);
            double lo = Double.POSITIVE_INFINITY, hi = 0;
            for (Node node : tree.getNodesAsArray()) {
                int i = node.getNr();
                height.setAtIndex(JAVA_DOUBLE, i, node.getHeight());
                parent.setAtIndex(JAVA_INT, i, node.isRoot() ? -1 : node.getParent().getNr());
                leafType.set(JAVA_BYTE, i, (byte) (i < nLeaves ? leafTypes[i] : 0));
                lo = Math.min(lo, node.getHeight());
                hi = Math.max(hi, node.getHeight());
            }
            // window: a lineage uses 1 + 2 nVirt uniforms, nVirt ~ Poisson(mu L)
            double muL = migrationModel.getMu(false) * (hi - lo);
            long win = 64 + (long) (nLeaves * (muL + 6 * Math.sqrt(muL + 1) + 4));
            for (;;) {
                MemorySegment order = a.allocate(4L * nParticles * nLeaves, 16),
                        u = a.allocate(8L * nParticles * win, 16), uOff = a.allocate(8L * (nParticles + 1), 8),
                        logPOut = a.allocate(8, 8), status = a.allocate(4L * nParticles, 4);
                for (int p = 0; p < nParticles; p++) {
                    int[] perm = Randomizer.shuffled(nLeaves);                          // :129
                    for (int i = 0; i < nLeaves; i++)
                        order.setAtIndex(JAVA_INT, (long) p * nLeaves + i, perm[i]);
                    for (long i = 0; i < win; i++)
                        u.setAtIndex(JAVA_DOUBLE, p * win + i, Randomizer.nextDouble());
                    uOff.setAtIndex(JAVA_LONG, p + 1, (p + 1) * win);
                }
                nativeLib.untypedLogPBatch(1, nParticles, MemorySegment.NULL, N, nodeOff, height, parent, leafType,
                        order, u, uOff, logPOut, MemorySegment.NULL, status, MemorySegment.NULL);
                boolean more = false;
                for (int p = 0; p < nParticles; p++)
                    more |= status.getAtIndex(JAVA_INT, p) == MttNative.ST_NEED_MORE_UNIFORMS;
                if (!more) {
                    logP = logPOut.get(JAVA_DOUBLE, 0);
                    return logP;
                }
                win *= 4;
            }
        }
    }

    @Override
    protected boolean requiresRecalculation() {
        return true;
    }

    @Override
    public boolean isStochastic() {
        return true;
    }
}
