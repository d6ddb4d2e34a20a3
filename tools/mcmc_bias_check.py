"""Long lock-step MCMC runs of the reference's operator-test configurations; reports the root
height mean / variance with a between-chain standard error, next to direct simulation.
usage: python tools/mcmc_bias_check.py [nChains] [steps] [config ...]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import multitypetree_b200 as mtt
from multitypetree_b200 import operators as ops
from multitypetree_b200.mcmc import Chain, MCMCEnsemble
from multitypetree_b200.uniformization import Randomizer

C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
which = sys.argv[3:] or ["STX", "NR_MTU_TS", "STX_TS", "TWB"]

m = np.array([[0.0, 0.1], [0.1, 0.0]])
N = np.array([7.0, 7.0])
CONFIGS = {
    "STX": ([1, 1, 0, 0], lambda t, mm, r: [ops.TypedSubtreeExchange(t, mm, randomizer=r),
                                            ops.NodeRetype(t, mm, randomizer=r),
                                            ops.MultiTypeUniform(t, mm, randomizer=r),
                                            ops.MultiTypeTreeScale(t, mm, scaleFactor=1.5, randomizer=r)]),
    "NR_MTU_TS": ([1, 1, 0, 0], lambda t, mm, r: [ops.NodeRetype(t, mm, randomizer=r),
                                                  ops.MultiTypeUniform(t, mm, randomizer=r),
                                                  ops.MultiTypeTreeScale(t, mm, scaleFactor=1.5, randomizer=r)]),
    "STX_TS": ([1, 1, 0, 0], lambda t, mm, r: [ops.TypedSubtreeExchange(t, mm, randomizer=r),
                                               ops.MultiTypeTreeScale(t, mm, scaleFactor=1.5, randomizer=r)]),
    "TWB": ([0, 0, 0], lambda t, mm, r: [ops.TypedWilsonBalding(t, mm, alpha=0.2, randomizer=r),
                                         ops.MultiTypeTreeScale(t, mm, scaleFactor=0.8, randomizer=r)]),
    "NSR": ([1, 0], lambda t, mm, r: [ops.NodeShiftRetype(t, mm, randomizer=r)]),
    "NSR4": ([1, 1, 0, 0], lambda t, mm, r: [ops.NodeShiftRetype(t, mm, randomizer=r),
                                             ops.TypedSubtreeExchange(t, mm, randomizer=r)]),
}
ctx = mtt.default_context()
for name in which:
    lt, mk = CONFIGS[name]
    n = len(lt)
    model = mtt.SCMigrationModel([0.1, 0.1], [7.0, 7.0], mtt.TypeSet("A", "B"), ctx=ctx)
    start = mtt.simulate_trees(C, n, 2, m, N, leafTypes=lt, leafTimes=np.zeros(n), seed=12345)
    chains = []
    for c in range(C):
        tree = start.tree(c)
        rng = Randomizer(1000 + c)
        chains.append(Chain(tree, model, mk(tree, model, rng), rng))
    ens = MCMCEnsemble(chains, ctx=ctx)
    t0 = time.time()
    ens.run(STEPS, logEvery=20)
    dt = time.time() - t0
    H = np.array(ens.heights)                   # chains start in equilibrium: no burn-in needed
    cm, cv = H.mean(axis=1), H.var(axis=1, ddof=1)
    sim = mtt.simulate_trees(400000, n, 2, m, N, leafTypes=lt, leafTimes=np.zeros(n), seed=77)
    hs = sim.height.reshape(-1, 2 * n - 1).max(axis=1)
    acc = sum(c.accepted for c in chains) / np.maximum(sum(c.proposed for c in chains), 1)
    print("%-10s mean %.3f +- %.3f (sim %.3f)  E[h^2] %.1f +- %.1f (sim %.1f)  acc %s  %.0f s"
          % (name, cm.mean(), cm.std(ddof=1) / np.sqrt(C), hs.mean(),
             (H ** 2).mean(), (H ** 2).mean(axis=1).std(ddof=1) / np.sqrt(C), (hs ** 2).mean(),
             np.round(acc, 3), dt), flush=True)
