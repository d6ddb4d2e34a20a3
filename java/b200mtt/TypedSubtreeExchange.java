/*
 * Drop-in replacement of multitypetree.operators.TypedSubtreeExchange
 * (operators/TypedSubtreeExchange.java:31-91 (narrow / wide subtree exchange, the two moved branches retyped)):
 * the reference's own operator -- same Inputs, same tree surgery, same Randomizer order -- with
 * the two uniformization methods it inherits from UniformizationRetypeOperator evaluated on the
 * device (DeviceBranchSampler: retypeBranch :116-225, getBranchTypeProb :233-275).
 * NOT compiled here (no JDK in the build container).
 */
package b200mtt;

import beast.base.core.Description;
import beast.base.evolution.tree.Node;
import multitypetree.evolution.tree.MultiTypeNode;

@Description("TypedSubtreeExchange with the branch typings sampled and scored on a B200.")
public class TypedSubtreeExchange extends multitypetree.operators.TypedSubtreeExchange {

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
