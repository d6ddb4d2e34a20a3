// C++ host-side mirror test: the reference's own known-answer test
// (test/multitypetree/coalescent/SCLikelihoodTest.java:43-73) written against
// include/mtt_b200.hpp, line for line:
//     MultiTypeTreeFromNewick mtTree ... "value", newickStr, "typeLabel", "state"
//     SCMigrationModel migrationModel ... "rateMatrix", "2.0 1.0", "popSizes", "5.0 10.0", typeSet {A, B}
//     StructuredCoalescentTreeDensity likelihood ... "migrationModel", "multiTypeTree"
//     assertEquals(-16.52831, likelihood.calculateLogP(), 1e-5)
// Usage: test_kat            -> runs on cuda:0, prints "KAT <logP>", exit 0 iff within 1e-5
//        test_kat --no-gpu   -> checks that, without a device, context creation fails loudly
#include <cmath>
#include <cstdio>
#include <cstring>

#include "mtt_b200.hpp"

int main(int argc, char **argv)
{
    const std::string newickStr =
        "(((A[&state=1]:0.25)[&state=0]:0.25,B[&state=0]:0.5)[&state=0]:1.5,"
        "(C[&state=0]:1.0,D[&state=0]:1.0)[&state=0]:1.0)[&state=0]:0.0;";
    mtt::MultiTypeTree mtTree = mtt::MultiTypeTree::fromNewick(newickStr, "state");
    if (mtTree.getLeafNodeCount() != 4 || mtTree.getTotalNumberOfChanges() != 1 ||
        mtTree.nodeType[0] != 1 || mtTree.changeTypes[0][0] != 0 ||
        std::fabs(mtTree.changeTimes[0][0] - 0.25) > 1e-15) {
        std::printf("newick parse FAILED\n");
        return 2;
    }
    mtt::TypeSet typeSet{"A", "B"};
    mtt::SCMigrationModel migrationModel({2.0, 1.0}, {5.0, 10.0}, typeSet);
    if (migrationModel.getBackwardRate(0, 1) != 2.0 || migrationModel.getBackwardRate(1, 0) != 1.0) {
        std::printf("rate resolution FAILED\n");
        return 2;
    }
    // the three rateMatrix encodings of SCMigrationModel.java:119-133
    mtt::TypeSet ts3{"a", "b", "c"};
    mtt::SCMigrationModel sq({0, 1, 2, 3, 0, 4, 5, 6, 0}, {1, 1, 1}, ts3), od({1, 2, 3, 4, 5, 6}, {1, 1, 1}, ts3),
        sy({1.5, 2.5, 3.5}, {1, 1, 1}, ts3);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            if (i != j && sq.getBackwardRate(i, j) != od.getBackwardRate(i, j)) {
                std::printf("encoding FAILED\n");
                return 2;
            }
    if (sy.getBackwardRate(2, 1) != 3.5 || sy.getBackwardRate(1, 2) != 3.5 || sy.getBackwardRate(0, 2) != 2.5) {
        std::printf("symmetric encoding FAILED\n");
        return 2;
    }

    if (argc > 1 && std::strcmp(argv[1], "--no-gpu") == 0) {
        try {
            mtt::Context ctx(0);
        } catch (const mtt::Error &e) {
            std::printf("no device: %s\n", e.what());
            return 0;          // failing loudly is the expected behaviour
        }
        std::printf("a device is present\n");
        return 0;
    }

    mtt::Context ctx(0);
    mtt::StructuredCoalescentTreeDensity likelihood(migrationModel, mtTree, /*checkValidity=*/true, ctx);
    const double expResult = -16.52831;      // "Calculated by hand" in the reference
    const double result = likelihood.calculateLogP();
    std::printf("KAT %.12f\n", result);
    // an invalid colouring gives -inf with checkValidity (StructuredCoalescentTreeDensity.java:90-91)
    mtt::MultiTypeTree bad = mtTree;
    bad.changeTypes[0][0] = 1;
    mtt::StructuredCoalescentTreeDensity badLik(migrationModel, bad, true, ctx);
    const double badResult = badLik.calculateLogP();
    std::printf("INVALID %f\n", badResult);
    return (std::fabs(result - expResult) < 1e-5 && std::isinf(badResult) && badResult < 0) ? 0 : 1;
}
