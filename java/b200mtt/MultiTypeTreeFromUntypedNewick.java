/*
 * Drop-in replacement of multitypetree.evolution.tree.MultiTypeTreeFromUntypedNewick
 * (evolution/tree/MultiTypeTreeFromUntypedNewick.java:32-372): same Inputs ("value",
 * "adjustTipHeights", "migrationModel" and, inherited from MultiTypeTree, "typeLabel" /
 * "typeTrait"), same result -- a typed tree with the Newick's topology, heights and leaf types
 * whose internal nodes and branches were coloured lineage by lineage (:104-134).  The colouring
 * itself (colourFirstLineage :150-199, colourLineage :202-212, addTypeChanges :272-357) is ONE
 * particle of mtt_untyped_sample_batch; this class flattens the tree, draws the leaf order
 * (Randomizer.shuffled, :116) and a window of uniforms from BEAST's Randomizer, and replays the
 * returned changes with MultiTypeNode.addChange.  A colouring without a valid path
 * (NoValidPathException, :128-131) comes back as a status and is retried with a new leaf order.
 * NOT compiled here (no JDK in the build container).
 */
package b200mtt;

import beast.base.core.Description;
import beast.base.core.Input;
import beast.base.core.Input.Validate;
import beast.base.core.Log;
import beast.base.evolution.tree.Node;
import beast.base.evolution.tree.Tree;
import beast.base.evolution.tree.TreeParser;
import beast.base.inference.StateNode;
import beast.base.inference.StateNodeInitialiser;
import beast.base.util.Randomizer;
import multitypetree.evolution.tree.MultiTypeNode;
import multitypetree.evolution.tree.MultiTypeTree;
import multitypetree.evolution.tree.SCMigrationModel;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.util.List;

import static java.lang.foreign.ValueLayout.*;

@Description("Class to initialize a MultiTypeTree from an untyped Newick tree (colouring on a B200).")
public class MultiTypeTreeFromUntypedNewick extends MultiTypeTree implements StateNodeInitialiser {

    public Input<String> newickStringInput = new Input<>("value", "Tree in Newick format.", Validate.REQUIRED);
    public Input<Boolean> adjustTipHeightsInput = new Input<>("adjustTipHeights",
            "Adjust tip heights in tree? Default true.", true);
    public Input<SCMigrationModel> migrationModelInput = new Input<>("migrationModel",
            "Migration model to use in simulator.", Validate.REQUIRED);
    public Input<Integer> deviceInput = new Input<>("device", "CUDA device ordinal (default: current).", -1);

    @Override
    public void initAndValidate() {
        super.initAndValidate();
        SCMigrationModel migrationModel = migrationModelInput.get();

        Tree flatTree = new TreeParser();
        flatTree.initByName("IsLabelledNewick", true, "adjustTipHeights", adjustTipHeightsInput.get(),
                "singlechild", true, "newick", newickStringInput.get());
        if (getTypeTrait() == null)
            throw new IllegalArgumentException(
                    "Trait set (with name '" + typeLabelInput.get() + "') must be provided.");   // :98-101

        // typed skeleton (:60-88): same numbering, heights, IDs, parents and child order
        final int n = flatTree.getNodeCount(), nLeaves = flatTree.getLeafNodeCount();
        MultiTypeNode[] typedNodes = new MultiTypeNode[n];
        for (int i = 0; i < n; i++) {
            typedNodes[i] = new MultiTypeNode();
            typedNodes[i].setNr(i);
            typedNodes[i].setHeight(flatTree.getNode(i).getHeight());
            typedNodes[i].setID(flatTree.getNode(i).getID());
        }
        for (int i = 0; i < n; i++) {
            Node node = flatTree.getNode(i);
            typedNodes[i].setParent(node.isRoot() ? null : typedNodes[node.getParent().getNr()]);
            for (int c = 0; c < node.getChildCount(); c++)
                typedNodes[i].getChildrenMutable().add(typedNodes[node.getChild(c).getNr()]);
        }
        for (int i = 0; i < nLeaves; i++)
            typedNodes[i].setNodeType(migrationModel.getTypeSet().getTypeIndex(
                    getTypeTrait().getStringValue(typedNodes[i].getID())));

        final int d = migrationModel.getNTypes();
        MttNative nativeLib = new MttNative(deviceInput.get());
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
            for (int i = 0; i < n; i++) {
                MultiTypeNode node = typedNodes[i];
                height.setAtIndex(JAVA_DOUBLE, i, node.getHeight());
                parent.setAtIndex(JAVA_INT, i, node.isRoot() ? -1 : node.getParent().getNr());
                leafType.set(JAVA_BYTE, i, (byte) (i < nLeaves ? node.getNodeType() : 0));
                lo = Math.min(lo, node.getHeight());
                hi = Math.max(hi, node.getHeight());
            }
            double muL = migrationModel.getMu(false) * (hi - lo);
            long win = 64 + (long) (nLeaves * (muL + 6 * Math.sqrt(muL + 1) + 4));
            int chgCap = Math.max(16, n);
            for (;;) {
                MemorySegment order = a.allocate(4L * nLeaves, 16), u = a.allocate(8L * win, 16),
                        uOff = a.allocate(16, 8), nodeType = a.allocate(n, 16), nChg = a.allocate(4, 4),
                        status = a.allocate(4, 4);
                int[] perm = Randomizer.shuffled(nLeaves);                               // :116
                for (int i = 0; i < nLeaves; i++)
                    order.setAtIndex(JAVA_INT, i, perm[i]);
                for (long i = 0; i < win; i++)
                    u.setAtIndex(JAVA_DOUBLE, i, Randomizer.nextDouble());
                uOff.setAtIndex(JAVA_LONG, 0, 0);
                uOff.setAtIndex(JAVA_LONG, 1, win);
                MemorySegment chgTime, chgType, chgNode;
                int st, c = 0;
                do {                            // the same draw again if the changes need more room
                    chgCap = Math.max(chgCap, 2 * c);
                    chgTime = a.allocate(8L * chgCap, 16);
                    chgType = a.allocate(chgCap, 16);
                    chgNode = a.allocate(4L * chgCap, 16);
                    nativeLib.untypedSampleBatch(1, 1, MemorySegment.NULL, N, nodeOff, height, parent, leafType,
                            order, u, uOff, chgCap, nodeType, nChg, chgTime, chgType, chgNode, MemorySegment.NULL,
                            status, MemorySegment.NULL);
                    st = status.get(JAVA_INT, 0);
                    c = nChg.get(JAVA_INT, 0);
                } while (st == 0 && c > chgCap);
                if (st == MttNative.ST_NEED_MORE_UNIFORMS) {
                    win *= 4;
                    continue;
                }
                if (st != 0) {
                    Log.info.println("Colour simulation failed. Retrying.");             // :129
                    continue;
                }
                for (int i = 0; i < n; i++) {
                    typedNodes[i].clearChanges();
                    typedNodes[i].setNodeType(nodeType.get(JAVA_BYTE, i) & 0xff);
                }
                for (int k = 0; k < c; k++)
                    typedNodes[chgNode.getAtIndex(JAVA_INT, k)].addChange(chgType.get(JAVA_BYTE, k) & 0xff,
                            chgTime.getAtIndex(JAVA_DOUBLE, k));
                break;
            }
        } finally {
            nativeLib.close();
        }

        // Construct MTT (:138-146)
        root = typedNodes[n - 1];
        root.setParent(null);
        nodeCount = root.getNodeCount();
        internalNodeCount = root.getInternalNodeCount();
        leafNodeCount = root.getLeafNodeCount();
        initArrays();
    }

    @Override
    public void initStateNodes() { }

    @Override
    public void getInitialisedStateNodes(List<StateNode> stateNodeList) {
        stateNodeList.add(this);
    }
}
