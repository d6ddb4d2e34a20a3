// simulator.cpp -- host-side generator of synthetic structured-coalescent multi-type trees.
//
// Benchmark / test INPUT generation only (never on the timed path).  It follows the semantics
// of the reference's Gillespie initialiser evolution/tree/StructuredCoalescentMultiTypeTree.java
// (simulateTree :191-289, updatePropensities :300-327, getNextEvent :352-385, updateTree
// :395-442) so that the synthetic trees have the shape the reference itself would start from
// (SURVEY.md section 8d), but writes them straight into the flattened tree-SoA layout of
// include/mtt_b200.h.  The RNG is xoshiro256** seeded per tree, NOT BEAST's Randomizer: the
// trees are reproducible across machines, not bit-identical to a BEAST run.
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <thread>
#include <vector>

namespace {

struct Rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t &x)
    {
        uint64_t z = (x += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed)
    {
        for (auto &v : s)
            v = splitmix(seed);
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next()
    {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return r;
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    int below(int n) { return (int)(uniform() * n); }
    double exponential(double rate) { return -std::log(1.0 - uniform()) / rate; }
};

struct SimTree {
    std::vector<double> height;
    std::vector<int32_t> parent;
    std::vector<uint8_t> nodeType;
    // type changes in simulation (= increasing time) order: (node, time, new type); a flat list
    // keeps the per-tree footprint small (2^20-tree batches are simulated for the benchmark)
    struct Change { int32_t node; double time; uint8_t type; };
    std::vector<Change> changes;
};

struct SimBatch {
    int64_t B;
    int nLeaves;
    std::vector<SimTree> trees;
};

void simulate_one(SimTree &t, int n, int d, const double *m, const double *N,
                  const uint8_t *leafTypes, const double *leafTimes, uint64_t seed)
{
    Rng rng(seed);
    const int nNodes = 2 * n - 1;
    t.height.assign(nNodes, 0.0);
    t.parent.assign(nNodes, -1);
    t.nodeType.assign(nNodes, 0);
    t.changes.clear();

    // current type of the lineage above each created node, per-deme active lists
    std::vector<std::vector<int>> active(d);
    std::vector<int> inactive(n);
    for (int i = 0; i < n; i++) {
        inactive[i] = i;
        t.height[i] = leafTimes[i];
        t.nodeType[i] = leafTypes[i];
    }
    std::stable_sort(inactive.begin(), inactive.end(),
                     [&](int a, int b) { return leafTimes[a] < leafTimes[b]; });
    size_t nextInactive = 0;
    int nextNr = n;
    int nActive = 0;
    double time = 0.0;
    std::vector<double> mu(d, 0.0);
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++)
            if (i != j)
                mu[i] += m[i * d + j];

    while (nActive > 1 || nextInactive < (size_t)n) {
        double total = 0.0;
        for (int i = 0; i < d; i++) {
            double k = (double)active[i].size();
            total += k * (k - 1) / (2.0 * N[i]) + k * mu[i];
        }
        double tEvent = INFINITY;
        if (total > 0.0)
            tEvent = time + rng.exponential(total);
        else if (nextInactive >= (size_t)n)
            break;   // disconnected demes and nothing left to activate: leave the forest as is
        double U = (total > 0.0) ? rng.uniform() * total : 0.0;

        if (nextInactive < (size_t)n && leafTimes[inactive[nextInactive]] < tEvent) {
            int leaf = inactive[nextInactive++];
            time = leafTimes[leaf];
            active[leafTypes[leaf]].push_back(leaf);
            nActive++;
            continue;
        }
        // choose the event (coalescence in i, or migration i -> j) in the reference's order
        bool done = false;
        for (int i = 0; i < d && !done; i++) {
            double k = (double)active[i].size();
            double pc = k * (k - 1) / (2.0 * N[i]);
            if (U < pc) {
                auto &L = active[i];
                int a = rng.below((int)L.size());
                int b = rng.below((int)L.size() - 1);
                if (b >= a)
                    b++;
                int daughter = L[a], son = L[b];
                int par = nextNr++;
                t.height[par] = tEvent;
                t.nodeType[par] = (uint8_t)i;
                t.parent[daughter] = par;
                t.parent[son] = par;
                L[a] = par;
                L.erase(L.begin() + b);
                nActive--;
                done = true;
                break;
            }
            U -= pc;
            for (int j = 0; j < d; j++) {
                if (j == i)
                    continue;
                double pm = k * m[i * d + j];
                if (U < pm) {
                    auto &L = active[i];
                    int a = rng.below((int)L.size());
                    int node = L[a];
                    t.changes.push_back({node, tEvent, (uint8_t)j});
                    L.erase(L.begin() + a);
                    active[j].push_back(node);
                    done = true;
                    break;
                }
                U -= pm;
            }
        }
        if (!done)
            continue;   // rounding fell through the propensity table: redraw
        time = tEvent;
    }
}

}  // namespace

extern "C" {

// Simulate B trees; leaf i has type leafTypes[i] and sampling time leafTimes[i].
// Tree b uses seed (seed + b).  Returns an opaque handle (NULL on bad arguments).
void *mtt_sim_run(int64_t B, int nLeaves, int d, const double *m, const double *N,
                  const uint8_t *leafTypes, const double *leafTimes, uint64_t seed, int nThreads)
{
    if (B < 0 || nLeaves < 2 || d < 1 || !m || !N || !leafTypes || !leafTimes)
        return nullptr;
    SimBatch *sb = new SimBatch();
    sb->B = B;
    sb->nLeaves = nLeaves;
    sb->trees.resize((size_t)B);
    if (nThreads < 1)
        nThreads = 1;
    auto work = [&](int tid) {
        for (int64_t b = tid; b < B; b += nThreads)
            simulate_one(sb->trees[(size_t)b], nLeaves, d, m, N, leafTypes, leafTimes,
                         seed + (uint64_t)b);
    };
    if (nThreads == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < nThreads; i++)
            th.emplace_back(work, i);
        for (auto &x : th)
            x.join();
    }
    return sb;
}

int64_t mtt_sim_total_changes(const void *h)
{
    const SimBatch *sb = (const SimBatch *)h;
    int64_t tot = 0;
    for (auto &t : sb->trees)
        tot += (int64_t)t.changes.size();
    return tot;
}

// Export into caller arrays: height/parent/nodeType/chgCount [B*(2n-1)], nodeOff/chgBase [B+1],
// chgTime/chgType [total changes].
void mtt_sim_export(const void *h, int64_t *nodeOff, double *height, int32_t *parent,
                    uint8_t *nodeType, uint16_t *chgCount, int64_t *chgBase, double *chgTime,
                    uint8_t *chgType, uint16_t *chgNode)
{
    const SimBatch *sb = (const SimBatch *)h;
    const int nNodes = 2 * sb->nLeaves - 1;
    int64_t c = 0;
    std::vector<int32_t> cnt(nNodes), pos(nNodes);
    for (int64_t b = 0; b < sb->B; b++) {
        const SimTree &t = sb->trees[(size_t)b];
        nodeOff[b] = b * nNodes;
        chgBase[b] = c;
        // counting sort of the change list by node; within a node the simulation order is
        // already youngest -> oldest
        std::fill(cnt.begin(), cnt.end(), 0);
        for (auto &ch : t.changes)
            cnt[ch.node]++;
        int32_t run = 0;
        for (int i = 0; i < nNodes; i++) {
            int64_t g = b * nNodes + i;
            height[g] = t.height[i];
            parent[g] = t.parent[i];
            nodeType[g] = t.nodeType[i];
            chgCount[g] = (uint16_t)cnt[i];
            pos[i] = run;
            run += cnt[i];
        }
        for (auto &ch : t.changes) {
            const int64_t k = c + pos[ch.node]++;
            chgTime[k] = ch.time;
            chgType[k] = ch.type;
            if (chgNode)
                chgNode[k] = (uint16_t)ch.node;        // owner layout of mtt_sc_logp_owner_batch
        }
        c += run;
    }
    nodeOff[sb->B] = sb->B * nNodes;
    chgBase[sb->B] = c;
}

void mtt_sim_free(void *h)
{
    delete (SimBatch *)h;
}

}  // extern "C"
