/*
 * TreeFlattener.java -- MultiTypeTree (object graph) -> tree SoA of include/mtt_b200.h.
 * Replaces pointer chasing over evolution/tree/MultiTypeNode.java:33-36 with one O(n + M) pass
 * into off-heap buffers that are reused between calls.  NOT compiled here (no JDK).
 */
package b200mtt;

import beast.base.evolution.tree.Node;
import multitypetree.evolution.tree.MultiTypeNode;
import multitypetree.evolution.tree.MultiTypeTree;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;

import static java.lang.foreign.ValueLayout.*;

final class TreeFlattener {
    final Arena arena = Arena.ofShared();
    MemorySegment nodeOff, height, parent, nodeType, chgNode, chgBase, chgTime, chgType;
    private int nodeCap = 0, chgCap = 0;

    /** Flatten one tree (B = 1).  Root-branch changes are not exported (density ignores them:
     *  StructuredCoalescentTreeDensity.java:161-162). */
    void flatten(MultiTypeTree tree) {
        int n = tree.getNodeCount();
        int M = 0;
        for (Node node : tree.getNodesAsArray())
            if (!node.isRoot())
                M += ((MultiTypeNode) node).getChangeCount();
        if (n > nodeCap) {
            nodeCap = n;
            height = arena.allocate(8L * n + 16, 16);
            parent = arena.allocate(4L * n + 16, 16);
            nodeType = arena.allocate(n + 16, 16);
            nodeOff = arena.allocate(16, 8);
            chgBase = arena.allocate(16, 8);
        }
        if (M > chgCap) {
            chgCap = Math.max(M, 2 * chgCap + 16);
            chgTime = arena.allocate(8L * chgCap + 16, 16);
            chgType = arena.allocate(chgCap + 16, 16);
            chgNode = arena.allocate(2L * chgCap + 16, 16);   // owner (node Nr) of every change
        }
        nodeOff.setAtIndex(JAVA_LONG, 0, 0);
        nodeOff.setAtIndex(JAVA_LONG, 1, n);
        int c = 0;
        for (int nr = 0; nr < n; nr++) {                 // Nr order; leaves are [0, nLeaves)
            MultiTypeNode node = (MultiTypeNode) tree.getNode(nr);
            height.setAtIndex(JAVA_DOUBLE, nr, node.getHeight());
            parent.setAtIndex(JAVA_INT, nr, node.isRoot() ? -1 : node.getParent().getNr());
            nodeType.set(JAVA_BYTE, nr, (byte) node.getNodeType());
            int k = node.isRoot() ? 0 : node.getChangeCount();
            for (int i = 0; i < k; i++) {                // youngest -> oldest
                chgTime.setAtIndex(JAVA_DOUBLE, c, node.getChangeTime(i));
                chgType.set(JAVA_BYTE, c, (byte) node.getChangeType(i));
                chgNode.setAtIndex(JAVA_SHORT, c, (short) nr);
                c++;
            }
        }
        chgBase.setAtIndex(JAVA_LONG, 0, 0);
        chgBase.setAtIndex(JAVA_LONG, 1, c);
    }
}
