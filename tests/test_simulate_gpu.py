"""GPU tests of the device-side tree simulator (csrc/k5_simulate.cu) through
the C-ABI: the same seed gives the host generator's tree (topology, types and change lists
identical, times to rounding: the device's log() is not the host's bit for bit), the trees are
valid inputs of the density entry straight from device memory, and the root-height statistics
match the analytic coalescent expectation the reference's own tests use."""
import numpy as np
import pytest

import multitypetree_b200 as mtt
from multitypetree_b200 import _lib
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,n,d", [(64, 2, 2), (200, 5, 2), (300, 30, 3), (128, 100, 2),
                                   (32, 200, 10), (6, 1000, 4), (40, 60, 16)])
def test_same_seed_same_tree_as_the_host_generator(ctx, B, n, d):
    m, N = mtt.synthetic_model(d)
    host = mtt.simulate_trees(B, n, d, m, N, seed=77 + n, nThreads=4)
    dev = mtt.DeviceTreeBatch(B, n, d, m, N, seed=77 + n, ctx=ctx)
    got = dev.to_host()
    assert dev.total_changes == len(host.chgTime)
    assert np.array_equal(got.nodeOff, host.nodeOff) and np.array_equal(got.chgBase, host.chgBase)
    assert np.array_equal(got.parent, host.parent)
    assert np.array_equal(got.nodeType, host.nodeType)
    assert np.array_equal(got.chgType, host.chgType)
    assert np.array_equal(got.chgNode, host.chgNode)
    assert np.array_equal(got.chgCount, host.chgCount)
    np.testing.assert_allclose(got.height, host.height, rtol=1e-12, atol=0)
    np.testing.assert_allclose(got.chgTime, host.chgTime, rtol=1e-12, atol=0)
    dev.close()


def test_device_batch_feeds_the_density_entry(ctx):
    """simulate -> logP without leaving the device, checked against the oracle on the export."""
    import torch
    B, n, d = 4096, 100, 2
    m, N = mtt.synthetic_model(d)
    dev = mtt.DeviceTreeBatch(B, n, d, m, N, seed=2000, ctx=ctx)
    p = dev.pointers()
    dm = torch.from_numpy(np.ascontiguousarray(m).ravel()).cuda()
    dN = torch.from_numpy(np.ascontiguousarray(N)).cuda()
    logP = torch.zeros(B, dtype=torch.float64, device="cuda")
    valid = torch.zeros(B, dtype=torch.uint8, device="cuda")
    L = _lib.lib()
    ctx.check(L.mtt_sc_logp_owner_batch_dev(
        ctx.handle, B, d, 1, dm.data_ptr(), dN.data_ptr(), p["nodeOff"], p["height"], p["parent"],
        p["nodeType"], p["chgNode"], p["chgBase"], p["chgTime"], p["chgType"], 1,
        logP.data_ptr(), valid.data_ptr()))
    ctx.synchronize()
    tb = dev.to_host()
    want = orc.sc_logp_batch(tb, m, N, checkValidity=True, nThreads=8)
    assert np.all(valid.cpu().numpy() == 1) and np.all(np.isfinite(want))
    np.testing.assert_allclose(logP.cpu().numpy(), want, rtol=1e-9)
    dev.close()


def test_root_height_statistics(ctx):
    """Two leaves in one deme of size N coalesce after Exp(1/N): mean N, variance N^2
    (the check of StructuredCoalescentMultiTypeTree's own main(), :486-546, on 10^5 trees)."""
    B, Npop = 100000, 3.0
    m = np.array([[0.0, 0.0], [0.0, 0.0]])
    dev = mtt.DeviceTreeBatch(B, 2, 2, m, np.array([Npop, 1.0]), leafTypes=[0, 0],
                              leafTimes=[0.0, 0.0], seed=5, ctx=ctx)
    h = dev.to_host().height.reshape(B, 3)[:, 2]
    se = Npop / np.sqrt(B)
    assert abs(h.mean() - Npop) < 4 * se
    assert abs(h.var() - Npop ** 2) < 4 * np.sqrt(8.0 / B) * Npop ** 2
    dev.close()
    # two demes with migration: every tree is valid and ends in one lineage
    m2, N2 = mtt.synthetic_model(2)
    dev = mtt.DeviceTreeBatch(2000, 8, 2, m2, N2, seed=6, ctx=ctx)
    tb = dev.to_host()
    assert np.all(mtt.tree_validity(tb, ctx))
    assert np.all((tb.parent.reshape(2000, 15) < 0).sum(1) == 1)
    dev.close()
