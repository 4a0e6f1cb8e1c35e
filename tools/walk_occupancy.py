import importlib, sys, numpy as np
sys.path.insert(0, '/root/repo')
pkg = importlib.import_module("procedural-universe_b200")
for n in (1 << 20, 1 << 24):
    p = pkg.seed_galaxy_host(n, 42, 1.0)
    sim = pkg.Sim(mode=pkg.MODE_BARNESHUT, theta=0.5)
    sim.init(p)
    st = sim.walk_stats()
    h = sim.walk_occupancy().astype(np.float64)
    it = h.sum()
    print("n", n, st, "warp iterations", it, "per warp", it / (n / 32))
    print(" lane-slots busy %.3f" % ((h * np.arange(33)).sum() / (32 * it)))
    c = np.cumsum(h) / it
    for k in (0, 1, 2, 4, 8, 16, 24, 31, 32):
        print("  <=%2d lanes: %.3f of iterations" % (k, c[k]))
    sim.close()
