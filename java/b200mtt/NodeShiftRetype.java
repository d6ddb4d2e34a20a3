/*
 * Drop-in replacement of multitypetree.operators.NodeShiftRetype
 * (operators/NodeShiftRetype.java:33-146 (root and non-root node shifts with retyping of the attached branches)):
 * the reference's own operator -- same Inputs, same tree surgery, same Randomizer order -- with
 * the two uniformization methods it inherits from UniformizationRetypeOperator evaluated on the
 * device (DeviceBranchSampler: retypeBranch :116-225, getBranchTypeProb :233-275).
 * NOT compiled here (no JDK in the build container).
 */
package b200mtt;

import beast.base.core.Description;
import beast.base.evolution.tree.Node;
import multitypetree.evolution.tree.MultiTypeNode;

@Description("NodeShiftRetype with the branch typings sampled and scored on a B200.")
public class NodeShiftRetype extends multitypetree.operators.NodeShiftRetype {

    private final DeviceBranchSampler device = new DeviceBranchSampler();

    @Override
    public void initAndValidate() {
        super.initAndValidate();
        device.init(migModel, useSymmetrizedRatesInput.get());
    }

    @Override
    protected double retypeBranch(Node srcNode) throws NoValidPathException {
        double logProb = device.retype((MultiTypeNode) srcNode);
        if (Double.isNaN(logProb))
            throw new NoValidPathException() { };
        return logProb;
    }

    @Override
    protected double getBranchTypeProb(Node srcNode) {
        return device.branchTypeProb((MultiTypeNode) srcNode);
    }
}
