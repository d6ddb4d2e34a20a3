"""The C++ host-side mirror (include/mtt_b200.hpp) compiles against the C-ABI with plain g++ and
reproduces the reference's known-answer test through it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_test_kat")


def _build():
    import multitypetree_b200 as mtt
    mtt.lib()
    libdir = os.path.join(ROOT, "multitypetree_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_kat.cpp"), "-o", EXE, "-L", libdir,
           "-l:libmtt_b200.so", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu():
    exe = _build()
    r = subprocess.run([exe, "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "no device" in r.stdout or "a device is present" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_reference_kat():
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    kat = float([l for l in r.stdout.splitlines() if l.startswith("KAT")][0].split()[1])
    assert abs(kat - (-16.52831)) < 1e-5
    assert "INVALID -inf" in r.stdout


PLUG = os.path.join(ROOT, "tests", "cpp", "_test_plugins")


def _build_plugins():
    import multitypetree_b200 as mtt
    from oracle import oracle as orc
    mtt.lib()
    orc.build()
    libdir = os.path.join(ROOT, "multitypetree_b200")
    odir = os.path.join(ROOT, "oracle")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_plugins.cpp"), "-o", PLUG, "-L", libdir,
           "-l:libmtt_b200.so", "-L", odir, "-l:libmtt_oracle.so",
           "-Wl,-rpath," + libdir, "-Wl,-rpath," + odir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return PLUG


def test_cpp_plugins_compile():
    """TreeLikelihood / NodeRetype / NodeRetypeEnsemble of include/mtt_b200.hpp build with plain
    g++ against the C-ABI (the oracle is linked as the test's checker only)."""
    _build_plugins()


@pytest.mark.gpu
def test_cpp_tree_likelihood_and_node_retype_match_python_mirror():
    """The C++ TreeLikelihood agrees with the oracle, and the C++ lock-step NodeRetype driver
    reproduces the Python mirror chain by chain from the same seeds: identical trees (number
    of changes, node and change types), change times and Hastings ratios to rounding."""
    import numpy as np
    import multitypetree_b200 as mtt
    from multitypetree_b200.operators import NodeRetype, run_proposals
    from multitypetree_b200.uniformization import Randomizer
    exe = _build_plugins()
    C, n, d, steps = 6, 40, 5, 25
    r = subprocess.run([exe, str(C), str(n), str(d), str(steps)], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    tl = [l for l in r.stdout.splitlines() if l.startswith("TREELIK")][0].split()
    assert abs(float(tl[1]) - float(tl[2])) <= 1e-9 * abs(float(tl[2]))
    got = {int(l.split()[1]): l.split()[2:] for l in r.stdout.splitlines() if l.startswith("CHAIN")}
    assert len(got) == C
    # the same ensemble through the Python mirror
    m, N = mtt.synthetic_model(d)
    tb = mtt.simulate_trees(C, n, d, m, N, seed=5000, nThreads=4)
    ctx = mtt.Context(0)
    model = mtt.SCMigrationModel(m, N, mtt.TypeSet(*[str(i) for i in range(d)]), ctx=ctx)
    trees = [tb.tree(c) for c in range(C)]
    ops = [NodeRetype(trees[c], model, randomizer=Randomizer(1000 + c), ctx=ctx) for c in range(C)]
    sumHR, nInf = np.zeros(C), np.zeros(C, int)
    for _ in range(steps):
        for c, hr in enumerate(run_proposals(ops)):
            if np.isinf(hr):
                nInf[c] += 1
            else:
                sumHR[c] += hr
    for c in range(C):
        t = trees[c]
        st = sum(sum(x) for x in t.changeTimes)
        sty = sum(sum(int(v) for v in x) for x in t.changeTypes) + 7 * int(np.sum(t.nodeType))
        g = got[c]
        assert int(g[1]) == nInf[c]
        assert int(g[2]) == t.getTotalNumberOfChanges()
        assert int(g[4]) == sty
        assert abs(float(g[3]) - st) <= 1e-9 * max(1.0, abs(st))
        assert abs(float(g[0]) - sumHR[c]) <= 1e-8 * max(1.0, abs(sumHR[c]))
    # untyped trees: the C++ colouring (checked against the oracle inside the executable) and the
    # particle density reproduce the Python mirror from the same seeds
    from multitypetree_b200.untyped import (StructuredCoalescentUntypedTreeDensity,
                                            colour_untyped_trees)
    ut = mtt.simulate_trees(1, n, d, m, N, seed=6000, nThreads=4).tree(0)
    col = [l for l in r.stdout.splitlines() if l.startswith("COLOUR")][0].split()
    ty = colour_untyped_trees(model, [ut], Randomizer(77), ctx=ctx)[0]
    assert int(col[1]) == 1 and int(col[2]) == ty.getTotalNumberOfChanges()
    assert abs(float(col[3]) - sum(sum(x) for x in ty.changeTimes)) <= 1e-9 * max(1.0, float(col[3]))
    lp = mtt.StructuredCoalescentTreeDensity(model, ty, checkValidity=True, ctx=ctx).calculateLogP()
    assert abs(float(col[4]) - lp) <= 1e-9 * abs(lp)
    und = [l for l in r.stdout.splitlines() if l.startswith("UNTYPED")][0].split()
    up = StructuredCoalescentUntypedTreeDensity(model, ut, nParticles=32, randomizer=Randomizer(78),
                                                ctx=ctx).calculateLogP()
    assert abs(float(und[1]) - up) <= 1e-9 * abs(up)
