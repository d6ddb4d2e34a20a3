"""Development probe: device-side tree simulation (mtt_sim_trees_create) against the host
generator, config-2 shape."""
import sys
import time
sys.path.insert(0, ".")
import multitypetree_b200 as mtt

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
d = int(sys.argv[3]) if len(sys.argv) > 3 else 2
m, N = mtt.synthetic_model(d)
ctx = mtt.Context(0)
mtt.DeviceTreeBatch(1024, n, d, m, N, seed=1, ctx=ctx).close()
t0 = time.perf_counter()
dev = mtt.DeviceTreeBatch(B, n, d, m, N, seed=2000, ctx=ctx)
t1 = time.perf_counter()
tb = mtt.simulate_trees(B, n, d, m, N, seed=2000, nThreads=16)
t2 = time.perf_counter()
print("device: %d trees (%d tips, %d demes, %d changes) in %.3f s = %.2f M trees/s;  host generator, 16 threads: %.3f s"
      % (B, n, d, dev.total_changes, t1 - t0, B / (t1 - t0) / 1e6, t2 - t1))
