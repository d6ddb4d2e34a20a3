"""A small Metropolis-Hastings driver in the shape of BEAST 2's ``MCMC`` ([EXT]), enough to run
the reference's operator tests (test/multitypetree/operators/*_Test.java) end to end through
the device path: proposals by operators.py (uniformization kernels), target density by
``mtt_sc_logp_batch`` (density kernel).

``MCMCEnsemble`` steps C independent chains in lock step, so that one step costs one batched
retype call, one batched branch-probability call and ONE density launch over C trees instead
of C x 3 tiny launches.  A single chain is the C = 1 case (``MCMC``).

With a ``TreeLikelihood`` the target is the posterior of the reference's config 1
(examples/structuredCoalescent.xml:560-572: structured-coalescent density + tree likelihood);
the likelihood term runs on resident chains (``PeelChains``): a proposal recomputes only the
partials above what it changed, accept / reject keeps or drops them -- what BEAST's
``TreeLikelihood`` does with dirty nodes and ``store`` / ``restore``.

Per step and chain, in BEAST's order: operator choice (one ``nextDouble``), ``proposal()``,
then -- unless the proposal returned -inf -- ``logAlpha = newLogP - oldLogP + logHR`` and one
more ``nextDouble`` only if ``logAlpha < 0``.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import numpy as np

from . import _lib
from .density import sc_logp_batch
from .operators import MultiTypeTreeOperator, run_proposals
from .tree import MultiTypeTree, TreeBatch
from .uniformization import Randomizer


class Chain:
    """One MCMC chain: a tree (the State), its migration model, its operators (all built on
    `tree` and on the chain's Randomizer) and the running log-density."""

    def __init__(self, tree: MultiTypeTree, migrationModel,
                 operators: Sequence[MultiTypeTreeOperator], randomizer: Randomizer,
                 checkValidity: bool = False):
        self.tree = tree
        self.migModel = migrationModel
        self.operators = list(operators)
        self.rng = randomizer
        self.checkValidity = bool(checkValidity)
        for op in self.operators:
            if op.mtTree is not tree or op.rng is not randomizer:
                raise ValueError("operators of a chain must share its tree and Randomizer")
        w = np.array([op.weight for op in self.operators], float)
        self._cum = np.cumsum(w / w.sum())
        self.logP = None
        self.accepted = np.zeros(len(self.operators), np.int64)
        self.proposed = np.zeros(len(self.operators), np.int64)

    def selectOperator(self) -> int:
        """OperatorSchedule.selectOperator ([EXT]): one uniform against cumulative weights."""
        k = int(np.searchsorted(self._cum, self.rng.nextDouble(), side="right"))
        return min(k, len(self.operators) - 1)


def effective_sample_size(x: np.ndarray, maxLag: int = 2000) -> float:
    """ESS from the integrated autocorrelation time, truncated at the first non-positive sum
    of two adjacent autocovariances (the estimator behind BEAST's ESS logger, [EXT])."""
    x = np.asarray(x, float)
    n = len(x)
    if n < 4:
        return float(n)
    xc = x - x.mean()
    maxLag = min(maxLag, n - 1)
    gam = np.array([np.dot(xc[:n - k], xc[k:]) / (n - k) for k in range(maxLag)])
    if gam[0] <= 0:
        return float(n)
    var_stat = gam[0]
    k = 1
    while k < maxLag:
        if k % 2 == 0 and gam[k - 1] + gam[k] <= 0:
            break
        var_stat += 2.0 * gam[k]
        k += 1
    act = var_stat / gam[0]
    return n / max(act, 1e-12)


class MCMCEnsemble:
    """C chains advanced in lock step against the structured-coalescent tree density."""

    def __init__(self, chains: Sequence[Chain], ctx=None, treeLikelihood=None,
                 incremental: bool = True):
        self.chains = list(chains)
        self.ctx = ctx or _lib.default_context()
        self.heights: List[List[float]] = [[] for _ in self.chains]
        # tree likelihood term (optional): resident chains, or a full re-evaluation of every
        # proposal (incremental=False; the two agree to rounding)
        self.treeLikelihood = treeLikelihood
        self._peel = None
        if treeLikelihood is not None and incremental:
            self._peel = treeLikelihood.chains(len(self.chains))
        self._evaluate(range(len(self.chains)), init=True)

    def _likelihood(self, trees) -> np.ndarray:
        """Tree likelihood of one tree per chain (all chains, in chain order)."""
        par = np.stack([np.asarray(t.parent, np.int32) for t in trees])
        hgt = np.stack([np.asarray(t.height, np.float64) for t in trees])
        if self._peel is not None:
            return self._peel.update(par, hgt)
        return self.treeLikelihood.calculateLogPBatch(par, hgt)

    # one density launch over the listed chains' current trees
    def _density(self, idx: Sequence[int]) -> np.ndarray:
        trees = [self.chains[i].tree for i in idx]
        models = [self.chains[i].migModel for i in idx]
        batch = TreeBatch.from_trees(trees)
        check = any(self.chains[i].checkValidity for i in idx)
        if all(mm is models[0] for mm in models):
            m, N = models[0].dense()
        else:
            dense = [mm.dense() for mm in models]
            m = np.stack([x[0] for x in dense])
            N = np.stack([x[1] for x in dense])
        return sc_logp_batch(batch, m, N, checkValidity=check, ctx=self.ctx)

    def _evaluate(self, idx, init=False):
        idx = list(idx)
        lp = self._density(idx)
        for i, v in zip(idx, lp):
            self.chains[i].logP = float(v)
        if self.treeLikelihood is not None:
            lik = self._likelihood([c.tree for c in self.chains])
            if self._peel is not None:
                self._peel.accept()
            for i in idx:
                self.chains[i].logP += float(lik[i])

    def step(self) -> None:
        chains = self.chains
        picks = [c.selectOperator() for c in chains]
        saved = [c.tree.copy() for c in chains]
        hrs = run_proposals([c.operators[k] for c, k in zip(chains, picks)])
        live = [i for i, hr in enumerate(hrs) if hr != float("-inf")]
        newLogP = dict(zip(live, self._density(live))) if live else {}
        if self.treeLikelihood is not None and live:
            # a failed proposal may have left its tree half edited: those chains show their
            # saved tree (nothing to recompute) and are rejected below
            alive = set(live)
            lik = self._likelihood([c.tree if i in alive else saved[i]
                                    for i, c in enumerate(chains)])
            for i in live:
                newLogP[i] = float(newLogP[i]) + float(lik[i])
        accepted = np.zeros(len(chains), np.uint8)
        for i, c in enumerate(chains):
            c.proposed[picks[i]] += 1
            accept = False
            if i in newLogP:
                logAlpha = float(newLogP[i]) - c.logP + hrs[i]
                accept = logAlpha >= 0 or c.rng.nextDouble() < math.exp(logAlpha)
            if accept:
                c.logP = float(newLogP[i])
                c.accepted[picks[i]] += 1
                accepted[i] = 1
            else:
                c.tree.assignFrom(saved[i])
        if self._peel is not None and live:
            self._peel.accept(accepted)

    def run(self, chainLength: int, logEvery: int = 100) -> None:
        for it in range(chainLength):
            self.step()
            if (it + 1) % logEvery == 0:
                for i, c in enumerate(self.chains):
                    self.heights[i].append(float(c.tree.height[c.tree.getRoot()]))

    # -- util/MultiTypeTreeStatLogger.java: root-height mean / variance / ESS after burn-in
    def height_stats(self, burninFrac: float = 0.1):
        """mean / variance / ESS over all chains, plus between-chain standard errors of the mean
        and of the variance (the chains are independent, so these need no autocorrelation model)"""
        kept = []
        ess = 0.0
        for h in self.heights:
            h = np.asarray(h[int(len(h) * burninFrac):])
            kept.append(h)
            ess += effective_sample_size(h)
        allh = np.concatenate(kept)
        C = len(kept)
        cm = np.array([h.mean() for h in kept])
        c2 = np.array([(h ** 2).mean() for h in kept])
        se_mean = float(cm.std(ddof=1) / np.sqrt(C)) if C > 1 else float("nan")
        # var = E[h^2] - mean^2: delta method on the chain-level moments
        g = c2 - 2.0 * allh.mean() * cm
        se_var = float(g.std(ddof=1) / np.sqrt(C)) if C > 1 else float("nan")
        return dict(mean=float(allh.mean()), var=float(allh.var(ddof=1)), ess=float(ess),
                    n=len(allh), se_mean=se_mean, se_var=se_var)


class MCMC(MCMCEnsemble):
    """Single chain (BEAST ``MCMC`` with one State, one Distribution, a list of operators)."""

    def __init__(self, tree, migrationModel, operators, randomizer, checkValidity=False, ctx=None,
                 treeLikelihood=None, incremental=True):
        super().__init__([Chain(tree, migrationModel, operators, randomizer, checkValidity)], ctx,
                         treeLikelihood, incremental)
