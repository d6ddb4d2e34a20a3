// mtt_b200.hpp -- header-only C++17 host-side mirror of the reference's plugin interface on top
// of the C-ABI (mtt_b200.h).  The reference is compiled (Java) code whose toolchain is absent
// from the build image, so this is the compiled-language host side: same class names, Input
// names, call order and error behaviour as
//   evolution/tree/TypeSet.java, evolution/tree/SCMigrationModel.java:38-66, 93-144, 252-375,
//   evolution/tree/MultiTypeTreeFromNewick.java:30-43 + MultiTypeTree.java:538-665,
//   distributions/StructuredCoalescentTreeDensity.java:43-147,
//   operators/UniformizationRetypeOperator.java:74-275 + operators/NodeRetype.java:31-63,
//   and BEAST 2's Alignment / HKY / SiteModel / TreeLikelihood [EXT] as the reference wires them
//   (examples/structuredCoalescent.xml:532-572).
// The Java shims in java/b200mtt/ do exactly what these classes do (flatten, call, unflatten).
// No arithmetic of the hot path lives here; it all runs behind the C-ABI on the GPU.
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "mtt_b200.h"

// host-side Randomizer stream that lives in the same shared object (csrc/host/randomizer.cpp)
extern "C" {
void *mtt_rng_new(int64_t seed);
void mtt_rng_free(void *h);
int mtt_rng_set_seed(void *h, int64_t seed);
double mtt_rng_next_double(void *h);
int32_t mtt_rng_next_int(void *h, int32_t n);
void mtt_rng_peek_doubles(const void *h, int64_t n, double *out);
void mtt_rng_skip_doubles(void *h, int64_t n);
}

namespace mtt {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

class Context {
public:
    explicit Context(int device = -1)
    {
        int rc = mtt_ctx_create(device, &h_);
        if (rc != MTT_OK)
            throw Error(rc, mtt_last_error(nullptr));
    }
    ~Context() { mtt_ctx_destroy(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    mtt_ctx *get() const { return h_; }
    void check(int rc) const
    {
        if (rc != MTT_OK)
            throw Error(rc, mtt_last_error(h_));
    }

private:
    mtt_ctx *h_ = nullptr;
};

// evolution/tree/TypeSet.java: sorted set of type names; getNTypes() = max(size, 2) (:59-61)
class TypeSet {
public:
    TypeSet(std::initializer_list<std::string> names) : TypeSet(std::vector<std::string>(names)) {}
    explicit TypeSet(std::vector<std::string> names) : names_(std::move(names))
    {
        std::sort(names_.begin(), names_.end());
        names_.erase(std::unique(names_.begin(), names_.end()), names_.end());
    }
    int getNTypes() const { return std::max<int>((int)names_.size(), 2); }
    int getTypeIndex(const std::string &n) const
    {
        auto it = std::find(names_.begin(), names_.end(), n);
        if (it == names_.end())
            throw std::invalid_argument("unknown type name " + n);
        return (int)(it - names_.begin());
    }

private:
    std::vector<std::string> names_;
};

// evolution/tree/SCMigrationModel.java (Input resolution only; Q/R/R^k live on the device)
class SCMigrationModel {
public:
    std::vector<double> rateMatrix, popSizes;
    double rateMatrixScaleFactor = 1.0, popSizesScaleFactor = 1.0;
    std::vector<uint8_t> rateMatrixFlags;     // empty = all rates used
    bool useForwardMigrationRateMatrix = false;
    int nTypes = 0;

    SCMigrationModel(std::vector<double> rates, std::vector<double> pops, const TypeSet &ts)
        : rateMatrix(std::move(rates)), popSizes(std::move(pops)), nTypes(ts.getNTypes())
    {
        initAndValidate();
    }
    void initAndValidate()                                                 // :93-144
    {
        const int d = nTypes, dim = (int)rateMatrix.size();
        if ((int)popSizes.size() != d)
            throw std::invalid_argument("popSizes dimension does not match the type set");
        if (dim == d * d) {
            square_ = true;
            symmetric_ = false;
        } else if (dim == d * (d - 1)) {
            square_ = symmetric_ = false;
        } else if (dim == d * (d - 1) / 2) {
            square_ = false;
            symmetric_ = true;
        } else {
            throw std::invalid_argument("Migration matrix has incorrect number of elements "
                                        "for given deme count.");         // :127-129
        }
        if (!rateMatrixFlags.empty() && (int)rateMatrixFlags.size() != dim)
            throw std::invalid_argument("Migration rate flags array does not have same number "
                                        "of elements as migration rate matrix.");
    }
    int getNTypes() const { return nTypes; }
    int getArrayOffset(int i, int j) const                                 // :324-345
    {
        if (i == j)
            throw std::logic_error("Programmer error: requested migration rate array offset "
                                   "for diagonal element of migration rate matrix.");
        if (square_)
            return i * nTypes + j;
        if (symmetric_)
            return j < i ? i * (i - 1) / 2 + j : j * (j - 1) / 2 + i;
        return i * (nTypes - 1) + (j > i ? j - 1 : j);
    }
    double getPopSize(int i) const { return popSizesScaleFactor * popSizes[i]; }   // :373-375
    double getBackwardRate(int i, int j) const                             // :252-268
    {
        if (i == j)
            return 0;
        if (useForwardMigrationRateMatrix)
            return getForwardRate(j, i) * getPopSize(j) / getPopSize(i);
        const int off = getArrayOffset(i, j);
        if (!rateMatrixFlags.empty() && !rateMatrixFlags[off])
            return 0.0;
        return rateMatrixScaleFactor * rateMatrix[off];
    }
    double getForwardRate(int i, int j) const                              // :285-303
    {
        if (i == j)
            return 0.0;
        if (useForwardMigrationRateMatrix) {
            const int off = getArrayOffset(i, j);
            if (!rateMatrixFlags.empty() && !rateMatrixFlags[off])
                return 0.0;
            return rateMatrixScaleFactor * rateMatrix[off];
        }
        return getBackwardRate(j, i) * getPopSize(j) / getPopSize(i);
    }
    // what crosses the C-ABI: dense backward rates (zero diagonal) and population sizes
    void dense(std::vector<double> &m, std::vector<double> &N) const
    {
        const int d = nTypes;
        m.assign((size_t)d * d, 0.0);
        N.resize(d);
        for (int i = 0; i < d; i++) {
            N[i] = getPopSize(i);
            for (int j = 0; j < d; j++)
                if (i != j)
                    m[(size_t)i * d + j] = getBackwardRate(i, j);
        }
    }

private:
    bool square_ = false, symmetric_ = false;
};

// Flattened MultiTypeTree (evolution/tree/MultiTypeNode.java:33-36 as SoA; nodes [0, nLeaves)
// are the leaves).
class MultiTypeTree {
public:
    std::vector<double> height;
    std::vector<int32_t> parent;
    std::vector<uint8_t> nodeType;
    std::vector<std::vector<double>> changeTimes;     // per node, youngest -> oldest
    std::vector<std::vector<uint8_t>> changeTypes;

    int getNodeCount() const { return (int)height.size(); }
    int getLeafNodeCount() const { return (getNodeCount() + 1) / 2; }
    int getInternalNodeCount() const { return getNodeCount() - getLeafNodeCount(); }
    bool isRoot(int node) const { return parent[node] < 0; }
    // children in BEAST's left / right order: the order of appearance in parent[] (for a tree
    // read from Newick: the order of the Newick string)
    int getLeft(int node) const { return children(node).first; }
    int getRight(int node) const { return children(node).second; }
    void clearChanges(int node)
    {
        changeTimes[node].clear();
        changeTypes[node].clear();
    }
    void addChange(int node, int newType, double time)
    {
        changeTimes[node].push_back(time);
        changeTypes[node].push_back((uint8_t)newType);
    }
    int getChangeCount(int node) const { return (int)changeTimes[node].size(); }
    int getFinalType(int node) const
    {
        return changeTypes[node].empty() ? nodeType[node] : changeTypes[node].back();
    }
    int getTotalNumberOfChanges() const
    {
        size_t n = 0;
        for (auto &c : changeTimes)
            n += c.size();
        return (int)n;
    }

    // MultiTypeTreeFromNewick (MultiTypeTreeFromNewick.java:30-43): flat Newick whose
    // single-child nodes are type changes (MultiTypeTree.java:538-665).  Integer-valued
    // `[&typeLabel=k]` annotations; leaves numbered in order of appearance, internal nodes in
    // post-order (root last).
    static MultiTypeTree fromNewick(const std::string &value, const std::string &typeLabel)
    {
        Parser p{value, typeLabel, 0};
        std::vector<Flat> flat;
        const int root = p.parse(flat);
        // heights back from the youngest tip
        std::vector<double> depth(flat.size(), 0.0);
        std::vector<int> order{root};
        double maxd = 0;
        for (size_t i = 0; i < order.size(); i++)
            for (int c : flat[order[i]].children) {
                depth[c] = depth[order[i]] + flat[c].length;
                maxd = std::max(maxd, depth[c]);
                order.push_back(c);
            }
        MultiTypeTree t;
        std::vector<Build> nodes;
        std::vector<int> leaves, internals;
        build(flat, depth, maxd, root, -1, nodes, leaves, internals);
        std::vector<int> perm(leaves);
        perm.insert(perm.end(), internals.begin(), internals.end());
        std::vector<int> newnr(nodes.size());
        for (size_t i = 0; i < perm.size(); i++)
            newnr[perm[i]] = (int)i;
        for (int old : perm) {
            const Build &b = nodes[old];
            t.height.push_back(b.height);
            t.parent.push_back(b.parent < 0 ? -1 : newnr[b.parent]);
            t.nodeType.push_back((uint8_t)b.type);
            t.changeTimes.push_back(b.times);
            t.changeTypes.push_back(b.types);
        }
        return t;
    }

private:
    std::pair<int, int> children(int node) const
    {
        if (kids_.size() != height.size()) {
            kids_.assign(height.size(), {-1, -1});
            for (int v = 0; v < (int)parent.size(); v++)
                if (parent[v] >= 0) {
                    auto &k = kids_[parent[v]];
                    (k.first < 0 ? k.first : k.second) = v;
                }
        }
        return kids_[node];
    }
    mutable std::vector<std::pair<int, int>> kids_;
    struct Flat {
        std::vector<int> children;
        int type = -1;
        double length = 0;
    };
    struct Build {
        double height;
        int parent, type;
        std::vector<double> times;
        std::vector<uint8_t> types;
    };
    struct Parser {
        const std::string &s;
        const std::string &label;
        size_t pos;
        int parse(std::vector<Flat> &out)
        {
            Flat f;
            if (s[pos] == '(') {
                pos++;
                for (;;) {
                    f.children.push_back(parse(out));
                    if (s[pos] == ',') {
                        pos++;
                        continue;
                    }
                    if (s[pos] == ')') {
                        pos++;
                        break;
                    }
                    throw std::invalid_argument("bad Newick");
                }
            }
            while (pos < s.size() && std::string(":,()[;").find(s[pos]) == std::string::npos)
                pos++;                                             // label
            while (pos < s.size() && (s[pos] == '[' || s[pos] == ':')) {
                if (s[pos] == '[') {
                    const size_t end = s.find(']', pos);
                    const std::string body = s.substr(pos + 1, end - pos - 1);
                    const size_t k = body.find(label + "=");
                    if (k != std::string::npos) {
                        size_t v = k + label.size() + 1;
                        if (body[v] == '"' || body[v] == '\'')
                            v++;
                        f.type = std::stoi(body.substr(v));
                    }
                    pos = end + 1;
                } else {
                    pos++;
                    size_t used = 0;
                    f.length = std::stod(s.substr(pos), &used);
                    pos += used;
                }
            }
            out.push_back(f);
            return (int)out.size() - 1;
        }
    };
    static int build(const std::vector<Flat> &flat, const std::vector<double> &depth, double maxd,
                     int fnode, int parentId, std::vector<Build> &nodes, std::vector<int> &leaves,
                     std::vector<int> &internals)
    {
        std::vector<double> tim;
        std::vector<uint8_t> col;
        while (flat[fnode].children.size() == 1) {
            col.push_back((uint8_t)flat[fnode].type);
            tim.push_back(maxd - depth[fnode]);
            fnode = flat[fnode].children[0];
        }
        std::reverse(col.begin(), col.end());
        std::reverse(tim.begin(), tim.end());
        const int my = (int)nodes.size();
        nodes.push_back(Build{maxd - depth[fnode], parentId, flat[fnode].type, tim, col});
        if (flat[fnode].children.empty()) {
            leaves.push_back(my);
        } else if (flat[fnode].children.size() == 2) {
            for (int c : flat[fnode].children)
                build(flat, depth, maxd, c, my, nodes, leaves, internals);
            internals.push_back(my);
        } else {
            throw std::invalid_argument("multifurcations are not supported");
        }
        return my;
    }
};

// Batch of flattened trees in the C-ABI layout.
struct TreeBatch {
    std::vector<int64_t> nodeOff{0}, chgBase{0};
    std::vector<double> height, chgTime;
    std::vector<int32_t> parent;
    std::vector<uint8_t> nodeType, chgType;
    std::vector<uint16_t> chgNode;                      // owner (tree-local node) of every change
    int64_t size() const { return (int64_t)nodeOff.size() - 1; }
    void add(const MultiTypeTree &t)
    {
        const int n = t.getNodeCount();
        for (int i = 0; i < n; i++) {
            height.push_back(t.height[i]);
            parent.push_back(t.parent[i]);
            nodeType.push_back(t.nodeType[i]);
            const bool root = t.parent[i] < 0;          // root-branch changes are not exported
            if (!root)
                for (size_t k = 0; k < t.changeTimes[i].size(); k++) {
                    chgTime.push_back(t.changeTimes[i][k]);
                    chgType.push_back(t.changeTypes[i][k]);
                    chgNode.push_back((uint16_t)i);
                }
        }
        nodeOff.push_back((int64_t)height.size());
        chgBase.push_back((int64_t)chgTime.size());
    }
};

// distributions/StructuredCoalescentTreeDensity.java: Inputs migrationModel, multiTypeTree,
// checkValidity; calculateLogP() (:86-147) evaluated on the device.
class StructuredCoalescentTreeDensity {
public:
    StructuredCoalescentTreeDensity(const SCMigrationModel &migrationModel,
                                    const MultiTypeTree &multiTypeTree, bool checkValidity,
                                    const Context &ctx)
        : model_(migrationModel), tree_(multiTypeTree), check_(checkValidity), ctx_(ctx)
    {
        int top = 0;                                       // type-set compatibility (:81-83)
        for (auto v : tree_.nodeType)
            top = std::max<int>(top, v);
        if (top >= model_.getNTypes())
            throw std::invalid_argument("Tree and migration model have incompatible type sets.");
    }
    double calculateLogP()
    {
        TreeBatch b;
        b.add(tree_);
        std::vector<double> out = calculateLogPBatch(b);
        return logP = out[0];
    }
    std::vector<double> calculateLogPBatch(const TreeBatch &b)
    {
        std::vector<double> m, N, out((size_t)b.size());
        model_.dense(m, N);
        std::vector<uint8_t> valid((size_t)b.size());
        static const double zero = 0.0;
        static const uint8_t zero8 = 0;
        static const uint16_t zero16 = 0;
        ctx_.check(mtt_sc_logp_owner_batch(
            ctx_.get(), b.size(), model_.getNTypes(), 1, m.data(), N.data(), b.nodeOff.data(),
            b.height.data(), check_ ? b.parent.data() : nullptr, b.nodeType.data(),
            b.chgNode.empty() ? &zero16 : b.chgNode.data(), b.chgBase.data(),
            b.chgTime.empty() ? &zero : b.chgTime.data(),
            b.chgType.empty() ? &zero8 : b.chgType.data(), check_ ? 1 : 0, out.data(),
            check_ ? valid.data() : nullptr));
        return out;
    }
    double logP = NAN;

private:
    const SCMigrationModel &model_;
    const MultiTypeTree &tree_;
    bool check_;
    const Context &ctx_;
};


// ------------------------------------------------------------------------------------------
// BEAST Randomizer stream (MersenneTwisterFast [EXT]) -- the generator of csrc/host/randomizer.cpp
class Randomizer {
public:
    explicit Randomizer(int64_t seed = 127) : h_(mtt_rng_new(seed))
    {
        if (!h_)
            throw std::invalid_argument("Non zero random seed required");
    }
    ~Randomizer() { mtt_rng_free(h_); }
    Randomizer(const Randomizer &) = delete;
    Randomizer &operator=(const Randomizer &) = delete;
    double nextDouble() { return mtt_rng_next_double(h_); }
    int nextInt(int n) { return mtt_rng_next_int(h_, n); }
    void peekDoubles(int64_t n, double *out) const { mtt_rng_peek_doubles(h_, n, out); }
    void skipDoubles(int64_t n) { mtt_rng_skip_doubles(h_, n); }
    // Poisson draw of the steady-state tail (UniformizationRetypeOperator.java:100-102);
    // distributionally exact, not stream-compatible with a JVM ([EXT], see uniformization.py)
    int nextPoisson(double lam)
    {
        if (lam < 40) {
            const double p = std::exp(-lam);
            double r = 1.0;
            for (int n = 0;; n++) {
                r *= nextDouble();
                if (r < p)
                    return n;
            }
        }
        const double u = nextDouble();
        int n = 0;
        double logp = -lam, cdf = std::exp(logp);
        while (cdf < u && n < 10 * lam + 100) {
            n++;
            logp += std::log(lam) - std::log((double)n);
            cdf += std::exp(logp);
        }
        return n;
    }

private:
    void *h_;
};

// ------------------------------------------------------------------------------------------
// piece 2: BEAST 2 Alignment / HKY / SiteModel / TreeLikelihood [EXT] over mtt_peel_*
inline int nucleotideState(char ch)
{
    switch (ch) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    default: return 4;                       // gap / ambiguity code: missing data
    }
}

// Alignment.calcPatterns: distinct columns (sorted) with integer weights
class Alignment {
public:
    Alignment(const std::vector<std::string> &taxa, const std::vector<std::string> &sequences)
        : taxa_(taxa)
    {
        const size_t nT = sequences.size(), nS = nT ? sequences[0].size() : 0;
        std::map<std::vector<int32_t>, int32_t> cols;
        for (size_t s = 0; s < nS; s++) {
            std::vector<int32_t> col(nT);
            for (size_t t = 0; t < nT; t++)
                col[t] = nucleotideState(sequences[t][s]);
            cols[col]++;
        }
        const size_t P = cols.size();
        patterns.assign(nT * P, 0);
        size_t p = 0;
        for (auto &kv : cols) {
            for (size_t t = 0; t < nT; t++)
                patterns[t * P + p] = kv.first[t];
            patternWeight.push_back(kv.second);
            p++;
        }
        siteCount_ = (int)nS;
    }
    int getTaxonCount() const { return (int)taxa_.size(); }
    int getPatternCount() const { return (int)patternWeight.size(); }
    int getSiteCount() const { return siteCount_; }
    std::vector<int32_t> patterns;           // [taxon][pattern]
    std::vector<int32_t> patternWeight;

private:
    std::vector<std::string> taxa_;
    int siteCount_ = 0;
};

struct HKY {                                 // Inputs kappa, frequencies
    double kappa = 2.0;
    double frequencies[4] = {0.25, 0.25, 0.25, 0.25};
};

// SiteModel: Inputs substModel, mutationRate, proportionInvariant (one rate category here; the
// discrete-gamma quantiles live in the Python mirror, likelihood.py)
struct SiteModel {
    HKY substModel;
    double mutationRate = 1.0;
    double proportionInvariant = 0.0;
};

class TreeLikelihood {                       // Inputs data, tree, siteModel, branchRateModel
public:
    TreeLikelihood(const Alignment &data, const SiteModel &siteModel, double clockRate,
                   const Context &ctx)
        : data_(data), site_(siteModel), ctx_(ctx)
    {
        ctx_.check(mtt_peel_create(ctx_.get(), data.getTaxonCount(), data.getPatternCount(), 1, &h_));
        ctx_.check(mtt_peel_set_tip_states(h_, data.patterns.data()));
        ctx_.check(mtt_peel_set_pattern_weights(h_, data.patternWeight.data()));
        const double pVar = 1.0 - site_.proportionInvariant;
        const double rate = site_.mutationRate / pVar, prop = pVar;
        ctx_.check(mtt_peel_set_site_model(h_, &rate, &prop, site_.substModel.frequencies,
                                           site_.proportionInvariant, clockRate));
        ctx_.check(mtt_peel_set_hky(h_, site_.substModel.kappa));
    }
    ~TreeLikelihood() { mtt_peel_destroy(h_); }
    TreeLikelihood(const TreeLikelihood &) = delete;
    TreeLikelihood &operator=(const TreeLikelihood &) = delete;
    double calculateLogP(const MultiTypeTree &tree)
    {
        if (tree.getLeafNodeCount() != data_.getTaxonCount())
            throw std::invalid_argument("The number of nodes in the tree does not match the "
                                        "number of sequences");
        ctx_.check(mtt_peel_logl_batch(h_, 1, tree.parent.data(), tree.height.data(), &logP));
        return logP;
    }
    double logP = NAN;
    mtt_peel *handle() const { return h_; }
    const Context &context() const { return ctx_; }
    int getTaxonCount() const { return data_.getTaxonCount(); }

private:
    const Alignment &data_;
    SiteModel site_;
    const Context &ctx_;
    mtt_peel *h_ = nullptr;
};

// What TreeLikelihood [EXT] does BETWEEN MCMC steps, for many chains at once: the partials of
// every internal node stay in HBM (BeerLikelihoodCore's two banks per node), a proposal
// recomputes only the nodes above what it changed (TreeLikelihood.traverse / Node.isDirty()),
// accept() keeps or drops it (store / restore).  One tree per chain, over the likelihood's
// alignment (mtt_peel_chains_*).
class PeelChains {
public:
    PeelChains(const TreeLikelihood &likelihood, int nChains)
        : ctx_(likelihood.context()), K_(nChains), n_(likelihood.getTaxonCount())
    {
        ctx_.check(mtt_peel_chains_create(likelihood.handle(), nChains, &s_));
    }
    ~PeelChains() { mtt_peel_chains_destroy(s_); }
    PeelChains(const PeelChains &) = delete;
    PeelChains &operator=(const PeelChains &) = delete;
    // the proposals' trees, one per chain -> their log-likelihoods
    std::vector<double> update(const std::vector<const MultiTypeTree *> &trees)
    {
        if ((int)trees.size() != K_)
            throw std::invalid_argument("PeelChains::update: one tree per chain");
        const size_t N = 2 * (size_t)n_ - 1;
        par_.resize(K_ * N);
        hgt_.resize(K_ * N);
        for (int k = 0; k < K_; k++) {
            if (trees[k]->getLeafNodeCount() != n_)
                throw std::invalid_argument("The number of nodes in the tree does not match the "
                                            "number of sequences");
            std::copy(trees[k]->parent.begin(), trees[k]->parent.end(), par_.begin() + k * N);
            std::copy(trees[k]->height.begin(), trees[k]->height.end(), hgt_.begin() + k * N);
        }
        std::vector<double> logL(K_);
        ctx_.check(mtt_peel_chains_update(s_, par_.data(), hgt_.data(), nullptr, nullptr, nullptr,
                                          nullptr, logL.data()));
        return logL;
    }
    // the proposals as lists of changed nodes: chain k names node[off[k] .. off[k+1]) with their
    // new parent / height per list entry; everything else stays as accepted
    std::vector<double> updateSparse(const std::vector<int64_t> &off,
                                     const std::vector<int32_t> &node,
                                     const std::vector<int32_t> &parent,
                                     const std::vector<double> &height)
    {
        if ((int)off.size() != K_ + 1 || node.size() != (size_t)off.back() ||
            parent.size() != node.size() || height.size() != node.size())
            throw std::invalid_argument("PeelChains::updateSparse: off[nChains+1], one entry per "
                                        "listed node");
        std::vector<double> logL(K_);
        ctx_.check(mtt_peel_chains_update_sparse(s_, off.data(), node.data(), parent.data(),
                                                 height.data(), nullptr, nullptr, nullptr,
                                                 logL.data()));
        return logL;
    }
    // keep[k] != 0: chain k's proposal becomes its accepted state; empty = all
    void accept(const std::vector<uint8_t> &keep = {})
    {
        if (!keep.empty() && (int)keep.size() != K_)
            throw std::invalid_argument("PeelChains::accept: one flag per chain");
        ctx_.check(mtt_peel_chains_accept(s_, keep.empty() ? nullptr : keep.data()));
    }
    long long lastEntries() const { return (long long)mtt_peel_chains_last_entries(s_); }

private:
    const Context &ctx_;
    int K_, n_;
    mtt_peel_chains *s_ = nullptr;
    std::vector<int32_t> par_;
    std::vector<double> hgt_;
};

// ------------------------------------------------------------------------------------------
// piece 3: operators/UniformizationRetypeOperator.java + operators/NodeRetype.java.
// One NodeRetype per chain (its own tree and Randomizer); `NodeRetypeEnsemble::proposal()`
// advances all chains by one proposal in lock step -- ONE mtt_unif_node_retype_batch per step
// for the whole ensemble (config 5: 64 chains), the jobs that need host action (longer uniform
// window, Poisson tail) being resubmitted through mtt_unif_retype_batch.
class NodeRetypeEnsemble {
public:
    struct Chain {
        MultiTypeTree *tree;
        Randomizer *rng;
        int model;                           // row of the model set installed on the ctx
    };
    // m: nModels dense d x d backward-rate matrices; useSymmetrizedRates as in the reference
    NodeRetypeEnsemble(const Context &ctx, int nModels, int d, const std::vector<double> &m,
                       bool useSymmetrizedRates = false, int uniformWindow = 128)
        : ctx_(ctx), d_(d), window_(uniformWindow)
    {
        ctx_.check(mtt_unif_set_models(ctx_.get(), nModels, d, m.data(), useSymmetrizedRates));
        mu_.resize(nModels);
        std::vector<double> Q(d * d), R(d * d);
        for (int k = 0; k < nModels; k++)
            ctx_.check(mtt_unif_get_matrices(ctx_.get(), k, Q.data(), R.data(), &mu_[k]));
    }
    void add(MultiTypeTree &tree, Randomizer &rng, int model = 0)
    {
        chains_.push_back(Chain{&tree, &rng, model});
    }
    size_t size() const { return chains_.size(); }
    // measurement: wall time spent inside the two device calls, and retype calls made
    double secondsInCalls[2] = {0.0, 0.0};
    long long deviceCalls = 0;
    long long resubmits[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // by MTT_ST_* of the branch that stopped a job

    // One NodeRetype.proposal() per chain; returns the log Hastings ratios
    // (-infinity where the reference returns Double.NEGATIVE_INFINITY, NodeRetype.java:57-58).
    std::vector<double> proposal()
    {
        const size_t C = chains_.size();
        std::vector<double> logHR(C, 0.0);
        std::vector<int> node(C);
        std::vector<BranchList> br(C);
        // ---- node choice (:36-37) and probability of the current typing (:40-44)
        B_.clear();
        std::vector<int64_t> off{0};
        std::vector<double> ct;
        std::vector<uint8_t> cty;
        for (size_t c = 0; c < C; c++) {
            MultiTypeTree &t = *chains_[c].tree;
            node[c] = t.getLeafNodeCount() + chains_[c].rng->nextInt(t.getInternalNodeCount());
            if (!t.isRoot(node[c]))
                br[c].push_back(node[c]);
            br[c].push_back(t.getLeft(node[c]));
            br[c].push_back(t.getRight(node[c]));
            for (int nd : br[c]) {
                pushBranch(t, nd, chains_[c].model);
                ct.insert(ct.end(), t.changeTimes[nd].begin(), t.changeTimes[nd].end());
                cty.insert(cty.end(), t.changeTypes[nd].begin(), t.changeTypes[nd].end());
                off.push_back((int64_t)ct.size());
            }
        }
        // The probability of the current typing draws no random numbers, so it is evaluated in
        // the same device call as the first round of retyping below (the old node types are
        // captured here, before the new ones are drawn).
        const Branches oldB = B_;
        std::vector<double> lp(oldB.t0.size());
        std::vector<int32_t> st(oldB.t0.size());
        ct.push_back(0.0);
        cty.push_back(0);
        // ---- new node type (:47-48), then the branches node -> left -> right (:52-56)
        for (size_t c = 0; c < C; c++)
            chains_[c].tree->nodeType[node[c]] = (uint8_t)chains_[c].rng->nextInt(d_);
        struct Job {
            size_t chain;
            size_t start = 0;
            int window;
            int cap = 32;                // changes returned per branch (grown when exceeded)
            int32_t forced[3] = {-1, -1, -1};
            bool failed = false;
        };
        std::vector<Job> pending;
        for (size_t c = 0; c < C; c++) {
            Job j;
            j.chain = c;
            // uniforms the job is expected to consume: per branch one for the event count and
            // two per virtual event (Poisson(mu L), plus six standard deviations) -- so that a
            // long branch does not cost the whole ensemble a second round
            const MultiTypeTree &t = *chains_[c].tree;
            double need = 0.0;
            for (int nd : br[c]) {
                const double muL = mu_[chains_[c].model] * (t.height[t.parent[nd]] - t.height[nd]);
                need += 2.0 * (muL + 6.0 * std::sqrt(muL) + 8.0) + 1.0;
            }
            j.window = std::max(window_, ((int)need + 31) / 32 * 32);
            // a branch has at most (window - 1) / 2 virtual events, hence changes: half the
            // window is always enough room for the changes returned (and keeps the call's
            // outputs small enough for the staged, two-copy transfer)
            j.cap = std::max(32, j.window / 2);
            pending.push_back(j);
        }
        std::vector<char> noPath(C, 0);
        bool firstRound = true;
        while (!pending.empty()) {
            B_.clear();
            std::vector<int64_t> brOff{0}, uOff{0};
            std::vector<int32_t> forced;
            std::vector<double> us;
            int outCap = 0;
            for (Job &j : pending) {
                MultiTypeTree &t = *chains_[j.chain].tree;
                for (size_t i = j.start; i < br[j.chain].size(); i++) {
                    pushBranch(t, br[j.chain][i], chains_[j.chain].model);
                    forced.push_back(j.forced[i]);
                }
                brOff.push_back((int64_t)B_.t0.size());
                us.resize(us.size() + j.window);
                chains_[j.chain].rng->peekDoubles(j.window, us.data() + us.size() - j.window);
                uOff.push_back((int64_t)us.size());
                outCap = std::max(outCap, j.cap);
            }
            const size_t nBr = B_.t0.size();
            std::vector<int32_t> nChg(nBr), status(nBr), nVirt(nBr);
            std::vector<double> oct(nBr * outCap), olp(nBr);
            std::vector<uint8_t> octy(nBr * outCap);
            std::vector<int64_t> used(pending.size());
            const auto tc1 = std::chrono::steady_clock::now();
            if (firstRound)
                ctx_.check(mtt_unif_node_retype_batch(
                    ctx_.get(), (int64_t)lp.size(), oldB.mdl.data(), oldB.t0.data(),
                    oldB.t1.data(), oldB.a.data(), oldB.b.data(), off.data(), ct.data(),
                    cty.data(), lp.data(), st.data(), (int64_t)pending.size(), brOff.data(),
                    B_.mdl.data(), B_.t0.data(), B_.t1.data(), B_.a.data(), B_.b.data(),
                    forced.data(), us.data(), uOff.data(), outCap, nChg.data(), oct.data(),
                    octy.data(), olp.data(), status.data(), nVirt.data(), used.data()));
            else
                ctx_.check(mtt_unif_retype_batch(
                    ctx_.get(), (int64_t)pending.size(), brOff.data(), B_.mdl.data(),
                    B_.t0.data(), B_.t1.data(), B_.a.data(), B_.b.data(), forced.data(),
                    us.data(), uOff.data(), outCap, nChg.data(), oct.data(), octy.data(),
                    olp.data(), status.data(), nVirt.data(), used.data()));
            secondsInCalls[1] += std::chrono::duration<double>(std::chrono::steady_clock::now() - tc1).count();
            deviceCalls++;
            if (firstRound) {
                size_t k = 0;
                for (size_t c = 0; c < C; c++)
                    for (size_t i = 0; i < br[c].size(); i++)
                        logHR[c] += lp[k++];
                firstRound = false;
            }
            std::vector<Job> next;
            for (size_t ji = 0; ji < pending.size(); ji++) {
                Job &j = pending[ji];
                const size_t lo = (size_t)brOff[ji], hi = (size_t)brOff[ji + 1];
                bool more = false;
                for (size_t x = lo; x < hi; x++)
                    more |= status[x] == MTT_ST_NEED_MORE_UNIFORMS;
                if (more) {                              // nothing consumed: a longer window
                    resubmits[MTT_ST_NEED_MORE_UNIFORMS]++;
                    j.window *= 4;
                    j.cap = std::max(j.cap, j.window / 2);
                    next.push_back(j);
                    continue;
                }
                bool overflow = false;                   // more changes than were returned
                for (size_t x = lo; x < hi; x++)
                    overflow |= status[x] == MTT_ST_OK && nChg[x] > outCap;
                if (overflow) {                          // nothing consumed either: same draws
                    j.cap = 4 * outCap;
                    next.push_back(j);
                    continue;
                }
                MultiTypeTree &t = *chains_[j.chain].tree;
                size_t kOk = hi - lo;
                for (size_t x = lo; x < hi; x++)
                    if (status[x] != MTT_ST_OK) {
                        kOk = x - lo;
                        break;
                    }
                for (size_t i = 0; i < kOk; i++) {       // branches retyped successfully
                    const int nd = br[j.chain][j.start + i];
                    t.clearChanges(nd);
                    for (int q = 0; q < nChg[lo + i]; q++)
                        t.addChange(nd, octy[(lo + i) * outCap + q], oct[(lo + i) * outCap + q]);
                    logHR[j.chain] -= olp[lo + i];
                }
                chains_[j.chain].rng->skipDoubles(used[ji]);
                if (kOk == hi - lo)
                    continue;                            // job done
                const int why = status[lo + kOk];
                resubmits[why & 7]++;
                if (why == MTT_ST_NO_VALID_PATH) {       // NoValidPathException (:137-140)
                    noPath[j.chain] = 1;
                    continue;
                }
                if (why == MTT_ST_NEEDS_POISSON_TAIL) {  // (:96-102)
                    const int thresh = nVirt[lo + kOk];
                    const double muL = mu_[chains_[j.chain].model] *
                                       (B_.t1[lo + kOk] - B_.t0[lo + kOk]);
                    int nv;
                    do
                        nv = chains_[j.chain].rng->nextPoisson(muL);
                    while (nv < thresh);
                    j.forced[j.start + kOk] = nv;
                    j.start += kOk;
                    next.push_back(j);
                    continue;
                }
                // Pba rubbish / forward-backward fell through: retypeBranch returned -Infinity
                // and the proposal carries on with the remaining branches (:141-144, :179-190)
                logHR[j.chain] = -std::numeric_limits<double>::infinity();
                j.failed = true;
                j.start += kOk + 1;
                if (j.start < br[j.chain].size())
                    next.push_back(j);
            }
            pending.swap(next);
        }
        for (size_t c = 0; c < C; c++)
            if (noPath[c])
                logHR[c] = -std::numeric_limits<double>::infinity();
        return logHR;
    }

private:
    // the <= 3 branches of a proposal (node, left, right), without a heap allocation per chain
    struct BranchList {
        int v[3];
        int n = 0;
        void push_back(int x) { v[n++] = x; }
        size_t size() const { return (size_t)n; }
        int operator[](size_t i) const { return v[i]; }
        const int *begin() const { return v; }
        const int *end() const { return v + n; }
    };
    struct Branches {
        std::vector<double> t0, t1;
        std::vector<uint8_t> a, b;
        std::vector<int32_t> mdl;
        void clear()
        {
            t0.clear();
            t1.clear();
            a.clear();
            b.clear();
            mdl.clear();
        }
    } B_;
    void pushBranch(const MultiTypeTree &t, int nd, int model)
    {
        const int p = t.parent[nd];
        B_.t0.push_back(t.height[nd]);
        B_.t1.push_back(t.height[p]);
        B_.a.push_back(t.nodeType[nd]);
        B_.b.push_back(t.nodeType[p]);
        B_.mdl.push_back(model);
    }
    const Context &ctx_;
    int d_, window_;
    std::vector<double> mu_;
    std::vector<Chain> chains_;
};

// operators/NodeRetype.java as a single-chain operator: Inputs multiTypeTree, migrationModel,
// useSymmetrizedRates
class NodeRetype {
public:
    NodeRetype(MultiTypeTree &multiTypeTree, const SCMigrationModel &migrationModel,
               Randomizer &randomizer, const Context &ctx, bool useSymmetrizedRates = false)
    {
        std::vector<double> m, N;
        migrationModel.dense(m, N);
        ens_.reset(new NodeRetypeEnsemble(ctx, 1, migrationModel.getNTypes(), m,
                                          useSymmetrizedRates));
        ens_->add(multiTypeTree, randomizer, 0);
    }
    double proposal() { return ens_->proposal()[0]; }

private:
    std::unique_ptr<NodeRetypeEnsemble> ens_;
};

// ------------------------------------------------------------------------------------------
// untyped trees -- the particle density and the typings themselves

// Randomizer.shuffled(n) [EXT]: identity, then element i swapped with element i + nextInt(n - i)
inline std::vector<int32_t> shuffled(Randomizer &rng, int n)
{
    std::vector<int32_t> a(n);
    for (int i = 0; i < n; i++)
        a[i] = i;
    for (int i = 0; i < n; i++)
        std::swap(a[i], a[i + rng.nextInt(n - i)]);
    return a;
}

// What both untyped entries read of a tree and of the chain's Randomizer: topology, heights and
// leaf types; per particle a leaf order and one window of uniforms.
class UntypedParticles {
public:
    UntypedParticles(const SCMigrationModel &migrationModel, const MultiTypeTree &tree,
                     const Context &ctx)
        : tree_(tree), ctx_(ctx), d_(migrationModel.getNTypes())
    {
        std::vector<double> m;
        migrationModel.dense(m, N_);
        ctx_.check(mtt_unif_set_models(ctx_.get(), 1, d_, m.data(), 0));
        std::vector<double> Q((size_t)d_ * d_), R((size_t)d_ * d_);
        double mu = 0;
        ctx_.check(mtt_unif_get_matrices(ctx_.get(), 0, Q.data(), R.data(), &mu));
        const auto mm = std::minmax_element(tree.height.begin(), tree.height.end());
        const double muL = mu * (*mm.second - *mm.first);
        window = 64 + (int64_t)(tree.getLeafNodeCount() * (muL + 6 * std::sqrt(muL + 1) + 4));
        nodeOff_ = {0, tree.getNodeCount()};
    }
    int64_t window;                   // uniforms per particle (a lineage uses 1 + 2 nVirt)
    void draw(Randomizer &rng, int nParticles)
    {
        const int nLeaves = tree_.getLeafNodeCount();
        order_.clear();
        u_.resize((size_t)nParticles * window);
        uOff_.assign(nParticles + 1, 0);
        for (int p = 0; p < nParticles; p++) {
            const std::vector<int32_t> perm = shuffled(rng, nLeaves);
            order_.insert(order_.end(), perm.begin(), perm.end());
            rng.peekDoubles(window, u_.data() + (size_t)p * window);
            rng.skipDoubles(window);
            uOff_[p + 1] = (int64_t)(p + 1) * window;
        }
    }
    const MultiTypeTree &tree_;
    const Context &ctx_;
    int d_;
    std::vector<double> N_, u_;
    std::vector<int64_t> nodeOff_, uOff_;
    std::vector<int32_t> order_;
};

// distributions/StructuredCoalescentUntypedTreeDensity.java: Inputs tree, migrationModel,
// nParticles; calculateLogP() (:116-207) = one call of mtt_untyped_logp_batch.
class StructuredCoalescentUntypedTreeDensity {
public:
    StructuredCoalescentUntypedTreeDensity(const SCMigrationModel &migrationModel,
                                           const MultiTypeTree &tree, int nParticles,
                                           Randomizer &randomizer, const Context &ctx)
        : in_(migrationModel, tree, ctx), nParticles_(nParticles), rng_(randomizer)
    {
    }
    double calculateLogP()
    {
        std::vector<int32_t> status(nParticles_);
        for (;;) {
            in_.draw(rng_, nParticles_);
            const MultiTypeTree &t = in_.tree_;
            in_.ctx_.check(mtt_untyped_logp_batch(
                in_.ctx_.get(), 1, nParticles_, nullptr, in_.N_.data(), in_.nodeOff_.data(),
                t.height.data(), t.parent.data(), t.nodeType.data(), in_.order_.data(),
                in_.u_.data(), in_.uOff_.data(), &logP, nullptr, status.data(), nullptr));
            if (std::find(status.begin(), status.end(), (int32_t)MTT_ST_NEED_MORE_UNIFORMS) ==
                status.end())
                return logP;
            in_.window *= 4;
        }
    }
    bool isStochastic() const { return true; }
    double logP = NAN;

private:
    UntypedParticles in_;
    int nParticles_;
    Randomizer &rng_;
};

// evolution/tree/MultiTypeTreeFromUntypedNewick.java:104-134: the types of the internal nodes
// and the type changes of an untyped tree (topology, heights, leaf types) drawn lineage by
// lineage -- one particle of mtt_untyped_sample_batch per attempt; a colouring without a valid
// path starts over with a new leaf order (:118-133).
inline MultiTypeTree colourUntypedTree(const MultiTypeTree &untyped,
                                       const SCMigrationModel &migrationModel,
                                       Randomizer &randomizer, const Context &ctx)
{
    UntypedParticles in(migrationModel, untyped, ctx);
    const int n = untyped.getNodeCount();
    int chgCap = std::max(16, n);
    for (int attempt = 0; attempt < 1000; attempt++) {
        in.draw(randomizer, 1);
        std::vector<uint8_t> nodeType(n), chgType;
        std::vector<double> chgTime;
        std::vector<int32_t> chgNode;
        int32_t nChg = 0, status = 0;
        do {                                   // the same draw again if the changes need more room
            chgCap = std::max(chgCap, 2 * nChg);
            chgType.resize(chgCap);
            chgTime.resize(chgCap);
            chgNode.resize(chgCap);
            ctx.check(mtt_untyped_sample_batch(
                ctx.get(), 1, 1, nullptr, in.N_.data(), in.nodeOff_.data(), untyped.height.data(),
                untyped.parent.data(), untyped.nodeType.data(), in.order_.data(), in.u_.data(),
                in.uOff_.data(), chgCap, nodeType.data(), &nChg, chgTime.data(), chgType.data(),
                chgNode.data(), nullptr, &status, nullptr));
        } while (status == MTT_ST_OK && nChg > chgCap);
        if (status == MTT_ST_NEED_MORE_UNIFORMS) {
            in.window *= 4;
            continue;
        }
        if (status != MTT_ST_OK)
            continue;                          // "Colour simulation failed. Retrying." (:129)
        MultiTypeTree out;
        out.height = untyped.height;
        out.parent = untyped.parent;
        out.nodeType = nodeType;
        out.changeTimes.assign(n, {});
        out.changeTypes.assign(n, {});
        for (int k = 0; k < nChg; k++)
            out.addChange(chgNode[k], chgType[k], chgTime[k]);
        return out;
    }
    throw Error(MTT_ERR_STATE, "colour simulation failed 1000 times");
}

}  // namespace mtt
