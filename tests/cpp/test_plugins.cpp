// C++ host-side mirror test for pieces 2 and 3 (include/mtt_b200.hpp):
//   TreeLikelihood   against the CPU oracle (BeerLikelihoodCore semantics) on a small alignment;
//   NodeRetype       many chains in lock step (NodeRetypeEnsemble) on simulated 10-deme trees,
//                    the configuration of BASELINE config 5.  Prints one line per chain
//                    ("CHAIN c sumLogHR nInf nChanges sumChangeTimes sumTypes") that the pytest
//                    wrapper compares with the Python mirror run from the same seeds, and the
//                    chain-steps per second of the native driver.
// Usage: test_plugins <chains> <tips> <demes> <steps> [--time [uniform window]]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "mtt_b200.hpp"
#include "../../oracle/mtt_oracle.h"

extern "C" {
void *mtt_sim_run(int64_t B, int nLeaves, int d, const double *m, const double *N,
                  const uint8_t *leafTypes, const double *leafTimes, uint64_t seed, int nThreads);
int64_t mtt_sim_total_changes(const void *h);
void mtt_sim_export(const void *h, int64_t *nodeOff, double *height, int32_t *parent,
                    uint8_t *nodeType, uint16_t *chgCount, int64_t *chgBase, double *chgTime,
                    uint8_t *chgType, uint16_t *chgNode);
void mtt_sim_free(void *h);
}

static std::vector<mtt::MultiTypeTree> simulate(int B, int n, int d, const std::vector<double> &m,
                                                const std::vector<double> &N, uint64_t seed)
{
    std::vector<uint8_t> lt(n);
    std::vector<double> ltm(n);
    for (int i = 0; i < n; i++) {
        lt[i] = (uint8_t)(i % d);
        ltm[i] = 0.5 * (i % 4);
    }
    void *h = mtt_sim_run(B, n, d, m.data(), N.data(), lt.data(), ltm.data(), seed, 4);
    const int nN = 2 * n - 1;
    const int64_t tot = mtt_sim_total_changes(h);
    std::vector<int64_t> nodeOff(B + 1), chgBase(B + 1);
    std::vector<double> height((size_t)B * nN), chgTime(tot + 1);
    std::vector<int32_t> parent((size_t)B * nN);
    std::vector<uint8_t> nodeType((size_t)B * nN), chgType(tot + 1);
    std::vector<uint16_t> chgCount((size_t)B * nN), owner(tot + 1);
    mtt_sim_export(h, nodeOff.data(), height.data(), parent.data(), nodeType.data(),
                   chgCount.data(), chgBase.data(), chgTime.data(), chgType.data(), owner.data());
    mtt_sim_free(h);
    std::vector<mtt::MultiTypeTree> out(B);
    for (int b = 0; b < B; b++) {
        mtt::MultiTypeTree &t = out[b];
        t.height.assign(height.begin() + (size_t)b * nN, height.begin() + (size_t)(b + 1) * nN);
        t.parent.assign(parent.begin() + (size_t)b * nN, parent.begin() + (size_t)(b + 1) * nN);
        t.nodeType.assign(nodeType.begin() + (size_t)b * nN, nodeType.begin() + (size_t)(b + 1) * nN);
        t.changeTimes.assign(nN, {});
        t.changeTypes.assign(nN, {});
        for (int64_t c = chgBase[b]; c < chgBase[b + 1]; c++) {
            t.changeTimes[owner[c]].push_back(chgTime[c]);
            t.changeTypes[owner[c]].push_back(chgType[c]);
        }
    }
    return out;
}

int main(int argc, char **argv)
{
    const int C = argc > 1 ? std::atoi(argv[1]) : 8, n = argc > 2 ? std::atoi(argv[2]) : 30,
              d = argc > 3 ? std::atoi(argv[3]) : 4, steps = argc > 4 ? std::atoi(argv[4]) : 20;
    const bool timing = argc > 5 && std::strcmp(argv[5], "--time") == 0;
    const int window = argc > 6 ? std::atoi(argv[6]) : 128;
    mtt::Context ctx(0);

    // ---------------------------------------------------------------- piece 2
    {
        const std::string nw = "((t1[&s=0]:0.3,t2[&s=0]:0.3)[&s=0]:0.5,((t3[&s=0]:0.2,t4[&s=0]:0.2)"
                               "[&s=0]:0.4,t5[&s=0]:0.6)[&s=0]:0.2)[&s=0]:0.0;";
        mtt::MultiTypeTree tree = mtt::MultiTypeTree::fromNewick(nw, "s");
        std::vector<std::string> taxa{"t1", "t2", "t3", "t4", "t5"};
        std::vector<std::string> seq{"ACGTACGTTTGACCA-AAGT", "ACGTACGATTGACCANAAGT",
                                     "ACCTACGTTTGTCCAGAAGT", "ACCTACGTTTGTCCAGAAGA",
                                     "GCGTACTTTTGACCAGAAGT"};
        mtt::Alignment data(taxa, seq);
        mtt::SiteModel sm;
        sm.substModel.kappa = 2.7;
        const double fr[4] = {0.1, 0.2, 0.3, 0.4};
        std::memcpy(sm.substModel.frequencies, fr, sizeof fr);
        sm.mutationRate = 0.8;
        sm.proportionInvariant = 0.15;
        mtt::TreeLikelihood lik(data, sm, /*clockRate=*/1.3, ctx);
        const double got = lik.calculateLogP(tree);
        const double rate = sm.mutationRate / (1.0 - sm.proportionInvariant),
                     prop = 1.0 - sm.proportionInvariant;
        const double want = orc_tree_likelihood(5, data.getPatternCount(), 1, data.patterns.data(),
                                                data.patternWeight.data(), tree.parent.data(),
                                                tree.height.data(), &rate, &prop, fr, 2.7, 1.3,
                                                0.15, 2, nullptr);
        std::printf("TREELIK %.12f %.12f patterns %d of %d sites\n", got, want,
                    data.getPatternCount(), data.getSiteCount());
        if (!(std::fabs(got - want) <= 1e-9 * std::fabs(want)))
            return 1;

        // resident chains: a height move on one of two chains, accept, reject
        mtt::PeelChains chains(lik, 2);
        mtt::MultiTypeTree a = tree, b = tree;
        std::vector<double> l0 = chains.update({&a, &b});
        chains.accept();
        int v = b.parent[0];                           // (t1, t2), below the root
        if (b.parent[v] < 0)
            v = b.parent[2];
        b.height[v] = 0.5 * (b.height[v] + b.height[b.parent[v]]);
        std::vector<double> l1 = chains.update({&a, &b});
        const long long entries = chains.lastEntries();
        const double wantB = lik.calculateLogP(b);
        chains.accept({1, 0});                         // chain 1 rejects: back to `tree`
        std::vector<double> l2 = chains.update({&a, &tree});
        // the same move of chain 1 said as a list of one node
        std::vector<double> l3 = chains.updateSparse({0, 0, 1}, {v}, {b.parent[v]}, {b.height[v]});
        if (l3[0] != l0[0] || l3[1] != l1[1])
            return 1;
        chains.accept({0, 0});
        if (chains.update({&a, &tree})[1] != l0[1] || chains.lastEntries() != 0)
            return 1;
        std::printf("RESIDENT %.12f %.12f moved %.12f full %.12f entries %lld then %lld\n", l0[0],
                    l0[1], l1[1], wantB, entries, chains.lastEntries());
        auto close = [](double x, double y) { return std::fabs(x - y) <= 1e-12 * std::fabs(y); };
        if (!close(l0[0], got) || l0[1] != l0[0] || l1[0] != l0[0] || l2[1] != l0[1] ||
            !close(l1[1], wantB) || entries < 1 || entries > 3 || chains.lastEntries() != 0)
            return 1;
    }

    // ---------------------------------------------------------------- piece 3
    std::vector<double> m((size_t)d * d, 0.0), N(d);
    for (int c = 0; c < d; c++) {
        N[c] = 1.0 + 0.25 * c;
        for (int cp = 0; cp < d; cp++)
            if (c != cp)
                m[(size_t)c * d + cp] = (1.0 + 0.1 * ((c * d + cp) % 3)) / (d - 1);
    }

    // ---------------------------------------------------------------- untyped trees (8f rank 1)
    if (!timing) {
        std::vector<std::string> names;
        for (int c = 0; c < d; c++)
            names.push_back(std::to_string(c));
        mtt::TypeSet ts(names);
        mtt::SCMigrationModel model(m, N, ts);
        const mtt::MultiTypeTree untyped = simulate(1, n, d, m, N, 6000)[0];
        // the typing: the C++ mirror against the oracle fed the same leaf order and uniforms
        mtt::Randomizer r1(77), r2(77);
        const mtt::MultiTypeTree typed = mtt::colourUntypedTree(untyped, model, r1, ctx);
        std::vector<double> Q((size_t)d * d), Qs((size_t)d * d), R((size_t)d * d), Rs((size_t)d * d);
        double mu = 0, mus = 0;
        orc_update_matrices(d, m.data(), Q.data(), Qs.data(), &mu, &mus, R.data(), Rs.data());
        orc_unif_model om{d, Q.data(), R.data(), mu, orc_rpow_new(d, R.data())};
        const std::vector<int32_t> order = mtt::shuffled(r2, n);
        std::vector<double> u(1 << 16);
        r2.peekDoubles((int64_t)u.size(), u.data());
        const int cap = 4096;
        std::vector<int32_t> oType(untyped.getNodeCount()), oNode(cap), oCty(cap);
        std::vector<double> oTime(cap);
        int32_t oChg = 0;
        double oW = 0;
        int64_t oUsed = 0;
        const int ost = orc_untyped_particle_tree(&om, N.data(), untyped.getNodeCount(),
                                                  untyped.parent.data(), untyped.height.data(),
                                                  untyped.nodeType.data(), order.data(), u.data(),
                                                  (int64_t)u.size(), &oW, &oUsed, oType.data(), cap,
                                                  &oChg, oNode.data(), oCty.data(), oTime.data());
        orc_rpow_free(om.rpow);
        mtt::MultiTypeTree want = untyped;
        want.changeTimes.assign(untyped.getNodeCount(), {});
        want.changeTypes.assign(untyped.getNodeCount(), {});
        for (int k = 0; k < oChg; k++)
            want.addChange(oNode[k], oCty[k], oTime[k]);
        bool same = ost == ORC_ST_OK && typed.changeTypes == want.changeTypes;
        for (int i = 0; same && i < untyped.getNodeCount(); i++)     // (first lineage: log())
            for (size_t k = 0; k < want.changeTimes[i].size(); k++)
                same = same && std::fabs(typed.changeTimes[i][k] - want.changeTimes[i][k]) <=
                                   1e-15 * std::fabs(want.changeTimes[i][k]);
        for (int i = 0; i < untyped.getNodeCount(); i++)
            same = same && typed.nodeType[i] == (uint8_t)oType[i];
        // ... and a valid typed tree for the density of piece 1
        mtt::StructuredCoalescentTreeDensity dens(model, typed, true, ctx);
        const double lpTyped = dens.calculateLogP();
        double st = 0;
        for (auto &c : typed.changeTimes)
            for (double x : c)
                st += x;
        std::printf("COLOUR %d %d %.12f %.12f\n", same ? 1 : 0, typed.getTotalNumberOfChanges(), st,
                    lpTyped);
        if (!same || !std::isfinite(lpTyped))
            return 1;
        mtt::Randomizer r3(78);
        mtt::StructuredCoalescentUntypedTreeDensity ud(model, untyped, 32, r3, ctx);
        std::printf("UNTYPED %.12f\n", ud.calculateLogP());
    }
    std::vector<mtt::MultiTypeTree> trees = simulate(C, n, d, m, N, 5000);
    std::vector<std::unique_ptr<mtt::Randomizer>> rngs;
    mtt::NodeRetypeEnsemble ens(ctx, 1, d, m, false, window);
    for (int c = 0; c < C; c++) {
        rngs.emplace_back(new mtt::Randomizer(1000 + c));
        ens.add(trees[c], *rngs[c], 0);
    }
    std::vector<double> sumHR(C, 0.0);
    std::vector<int> nInf(C, 0);
    const auto t0 = std::chrono::steady_clock::now();
    for (int s = 0; s < steps; s++) {
        std::vector<double> hr = ens.proposal();
        for (int c = 0; c < C; c++) {
            if (std::isinf(hr[c]))
                nInf[c]++;
            else
                sumHR[c] += hr[c];
        }
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!timing)
        for (int c = 0; c < C; c++) {
            double st = 0;
            long sty = 0;
            for (int i = 0; i < trees[c].getNodeCount(); i++) {
                for (double x : trees[c].changeTimes[i])
                    st += x;
                for (uint8_t x : trees[c].changeTypes[i])
                    sty += x;
                sty += 7 * trees[c].nodeType[i];
            }
            std::printf("CHAIN %d %.12f %d %d %.12f %ld\n", c, sumHR[c], nInf[c],
                        trees[c].getTotalNumberOfChanges(), st, sty);
        }
    std::printf("CALLS branch_logp %.1f us/step, retype %.1f us/step (%.2f calls/step)\n",
                1e6 * ens.secondsInCalls[0] / steps, 1e6 * ens.secondsInCalls[1] / steps,
                (double)ens.deviceCalls / steps);
    std::printf("STOPS no-path %lld rubbish %lld fell-through %lld poisson-tail %lld more-uniforms %lld\n",
                ens.resubmits[MTT_ST_NO_VALID_PATH], ens.resubmits[MTT_ST_PBA_RUBBISH],
                ens.resubmits[MTT_ST_FB_FELL_THROUGH], ens.resubmits[MTT_ST_NEEDS_POISSON_TAIL],
                ens.resubmits[MTT_ST_NEED_MORE_UNIFORMS]);
    std::printf("RATE %d chains x %d steps in %.3f s = %.0f chain-steps/s (%.2f us per chain-step)\n",
                C, steps, secs, C * steps / secs, 1e6 * secs / (C * (double)steps));
    return 0;
}
