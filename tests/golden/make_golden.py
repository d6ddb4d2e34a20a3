"""Regenerates tests/golden/*.npz.  Run in the BUILD container (needs /root/reference for the
two NEXUS fixtures); the .npz files are committed because /root/reference does not exist on
the GPU box.

  density_fixtures.npz   flattened trees + expected logP
      kat          the reference's own known-answer test (SCLikelihoodTest.java:43-73):
                   expected value -16.52831 (tol 1e-5) comes from the REFERENCE, not from us
      true128      XML/simulated_data/true_tree.nexus (128 tips, 2 demes, 23 migrations)
      true64       XML/badness_eradication/true_tree.nexus (64 tips, 2 demes, 36 migrations)
      ewing3       inline 3-tip tree of test/multitypetree/operators/Ewing_Test.java:53-55
                   expected values for these three: oracle (faithful variant), cross-checked in
                   tests/test_oracle.py against the independent sufficient-statistics formula
                   and an mpmath re-evaluation.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import multitypetree_b200 as mtt  # noqa: E402
from oracle import oracle as orc  # noqa: E402

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def pack(prefix, tree, m, N, expected, out):
    tb = mtt.TreeBatch.from_trees([tree])
    for f in ("nodeOff", "height", "parent", "nodeType", "chgCount", "chgBase", "chgTime",
              "chgType"):
        out[prefix + "_" + f] = getattr(tb, f)
    out[prefix + "_m"] = m
    out[prefix + "_N"] = N
    out[prefix + "_expected"] = np.float64(expected)


def main():
    out = {}
    kat = ("(((A[&state=1]:0.25)[&state=0]:0.25,B[&state=0]:0.5)[&state=0]:1.5,"
           "(C[&state=0]:1.0,D[&state=0]:1.0)[&state=0]:1.0)[&state=0]:0.0;")
    model = mtt.SCMigrationModel([2.0, 1.0], [5.0, 10.0], mtt.TypeSet("A", "B"))
    m, N = model.dense()
    pack("kat", mtt.MultiTypeTreeFromNewick(kat, "state"), m, N, -16.52831, out)
    for name, path in (("true128", "XML/simulated_data/true_tree.nexus"),
                       ("true64", "XML/badness_eradication/true_tree.nexus")):
        t = mtt.MultiTypeTreeFromNewick(mtt.read_nexus_tree(os.path.join(REF, path)), "location")
        tb = mtt.TreeBatch.from_trees([t])
        pack(name, t, m, N, orc.sc_logp_batch(tb, m, N, variant=orc.FAITHFUL)[0], out)
    ewing = "((1[&deme=0]:1,2[&deme=0]:1)[&deme=0]:1,3[&deme=0]:2)[&deme=0]:0;"
    m3 = np.array([[0, .1, .2], [.3, 0, .4], [.5, .6, 0]])
    N3 = np.array([7.0, 7.0, 7.0])
    t = mtt.MultiTypeTreeFromNewick(ewing, "deme")
    pack("ewing3", t, m3, N3,
         orc.sc_logp_batch(mtt.TreeBatch.from_trees([t]), m3, N3, variant=orc.FAITHFUL)[0], out)
    np.savez_compressed(os.path.join(HERE, "density_fixtures.npz"), **out)
    print({k: float(v) for k, v in out.items() if k.endswith("_expected")})


if __name__ == "__main__":
    main()
