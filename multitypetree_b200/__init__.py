"""multitypetree_b200 -- B200-native likelihood hot path of the BEAST 2 MultiTypeTree package.

The product is the CUDA library ``libmtt_b200.so`` behind the C-ABI of ``include/mtt_b200.h``;
this package is the thin host-side mirror of the reference's plugin interface used by the
tests, the benchmark and as a model for the Java shims in ``java/``.
"""
from ._lib import Context, MttError, default_context, lib  # noqa: F401
from .tree import (DeviceTreeBatch, MultiTypeTree, MultiTypeTreeFromNewick, TreeBatch,  # noqa: F401
                   read_nexus_tree, read_nexus_trees, simulate_trees,
                   synthetic_model)
from .migration_model import SCMigrationModel, TypeSet  # noqa: F401
from .density import (StructuredCoalescentTreeDensity, sc_logp_batch,  # noqa: F401
                      sc_logp_tips_batch, set_tips, tree_validity)
from .uniformization import Randomizer, UniformizationModels  # noqa: F401
from .operators import (MultiTypeTreeScale, MultiTypeUniform, NodeRetype,  # noqa: F401
                        NodeShiftRetype, TypedSubtreeExchange, TypedWilsonBalding,
                        run_proposals)
from .mcmc import MCMC, Chain, MCMCEnsemble  # noqa: F401
from .untyped import (MultiTypeTreeFromUntypedNewick, StructuredCoalescentUntypedTree,  # noqa: F401
                      StructuredCoalescentUntypedTreeDensity, colour_untyped_trees,
                      untyped_sample_batch)

__all__ = [
    "Context", "MttError", "default_context", "lib", "DeviceTreeBatch", "MultiTypeTree",
    "MultiTypeTreeFromNewick", "TreeBatch", "read_nexus_tree", "read_nexus_trees", "simulate_trees",
    "synthetic_model", "SCMigrationModel", "TypeSet", "StructuredCoalescentTreeDensity",
    "sc_logp_batch", "sc_logp_tips_batch", "set_tips", "tree_validity", "Randomizer", "UniformizationModels", "NodeRetype",
    "NodeShiftRetype", "TypedSubtreeExchange", "TypedWilsonBalding", "MultiTypeTreeScale",
    "MultiTypeUniform", "run_proposals", "MCMC", "Chain", "MCMCEnsemble",
    "MultiTypeTreeFromUntypedNewick", "StructuredCoalescentUntypedTree",
    "StructuredCoalescentUntypedTreeDensity", "colour_untyped_trees", "untyped_sample_batch",
]
