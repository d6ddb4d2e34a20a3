/*
 * Drop-in replacement of multitypetree.distributions.StructuredCoalescentTreeDensity
 * (distributions/StructuredCoalescentTreeDensity.java:41-253): same Inputs ("multiTypeTree" from
 * MultiTypeTreeDistribution.java:35-36, "migrationModel", "checkValidity"), same contract of
 * calculateLogP(); the event sequence, validity check and interval scan run on the B200 through
 * mtt_sc_logp_owner_batch.  Register in version.xml as
 *   <provider classname="b200mtt.StructuredCoalescentTreeDensity"/>
 * and use spec="b200mtt.StructuredCoalescentTreeDensity" in the XML.  NOT compiled here (no JDK).
 */
package b200mtt;

import beast.base.core.Description;
import beast.base.core.Input;
import beast.base.core.Input.Validate;
import multitypetree.distributions.MultiTypeTreeDistribution;
import multitypetree.evolution.tree.MultiTypeTree;
import multitypetree.evolution.tree.SCMigrationModel;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;

import static java.lang.foreign.ValueLayout.*;

@Description("Likelihood of ColouredTree under structured coalescent (B200-native).")
public class StructuredCoalescentTreeDensity extends MultiTypeTreeDistribution {

    public Input<SCMigrationModel> migrationModelInput = new Input<>(
            "migrationModel", "Model of migration between demes.", Validate.REQUIRED);
    public Input<Boolean> checkValidityInput = new Input<>(
            "checkValidity", "Explicitly check validity of colouring.  (Default false.)", false);
    public Input<Integer> deviceInput = new Input<>("device", "CUDA device ordinal (default: current).", -1);

    protected SCMigrationModel migrationModel;
    protected MultiTypeTree mtTree;
    protected boolean checkValidity;

    private MttNative nativeLib;
    private TreeFlattener flat;
    private Arena arena;
    private MemorySegment m, N, logPOut, validOut;

    @Override
    public void initAndValidate() {
        migrationModel = migrationModelInput.get();
        mtTree = mtTreeInput.get();
        checkValidity = checkValidityInput.get();
        // same compatibility check as the reference (:81-83)
        if (mtTree.hasTypeTrait() && !mtTree.getTypeSet().equals(migrationModel.getTypeSet()))
            throw new IllegalArgumentException("Tree and migration model have incompatible type sets.");
        nativeLib = new MttNative(deviceInput.get());
        flat = new TreeFlattener();
        arena = Arena.ofShared();
        int d = migrationModel.getNTypes();
        m = arena.allocate(8L * d * d, 16);
        N = arena.allocate(8L * d, 16);
        logPOut = arena.allocate(8, 8);
        validOut = arena.allocate(1, 1);
    }

    @Override
    public double calculateLogP() {
        int d = migrationModel.getNTypes();
        // rates resolved by the reference's own accessors (all three encodings, scale factors,
        // BSSVS flags, forward-rate mode: SCMigrationModel.java:252-268, 373-375)
        for (int i = 0; i < d; i++) {
            N.setAtIndex(JAVA_DOUBLE, i, migrationModel.getPopSize(i));
            for (int j = 0; j < d; j++)
                m.setAtIndex(JAVA_DOUBLE, (long) i * d + j, i == j ? 0.0 : migrationModel.getBackwardRate(i, j));
        }
        flat.flatten(mtTree);
        nativeLib.scLogPBatch(1, d, 1, m, N, flat.nodeOff, flat.height,
                checkValidity ? flat.parent : MemorySegment.NULL, flat.nodeType, flat.chgNode,
                flat.chgBase, flat.chgTime, flat.chgType, checkValidity, logPOut,
                checkValidity ? validOut : MemorySegment.NULL);
        logP = logPOut.get(JAVA_DOUBLE, 0);      // -inf when checkValidity && !isValid (:90-91)
        return logP;
    }

    @Override
    public boolean requiresRecalculation() {
        return true;                             // as the reference (:250-253)
    }
}
