// mtt_b200.hpp -- header-only C++17 host-side mirror of the reference's plugin interface on top
// of the C-ABI (mtt_b200.h).  The reference is compiled (Java) code whose toolchain is absent
// from the build image, so this is the compiled-language host side: same class names, Input
// names, call order and error behaviour as
//   evolution/tree/TypeSet.java, evolution/tree/SCMigrationModel.java:38-66, 93-144, 252-375,
//   evolution/tree/MultiTypeTreeFromNewick.java:30-43 + MultiTypeTree.java:538-665,
//   distributions/StructuredCoalescentTreeDensity.java:43-147.
// The Java shims in java/b200mtt/ do exactly what these classes do (flatten, call, unflatten).
// No arithmetic of the hot path lives here; it all runs behind the C-ABI on the GPU.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "mtt_b200.h"

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
    TypeSet(std::initializer_list<std::string> names) : names_(names)
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

}  // namespace mtt
