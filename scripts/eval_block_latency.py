import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import sympleints_b200
eng = sympleints_b200.get_engine(0, 4, 5)
rng = np.random.default_rng(0)
sa = (np.array([1.2, 0.5]), np.array([0.6, 0.5]), rng.uniform(-1, 1, 3))
sb = (np.array([0.9]), np.array([1.0]), rng.uniform(-1, 1, 3))
sc = (np.array([0.7]), np.array([1.0]), rng.uniform(-1, 1, 3))
for key, Ls, args in (("ovlp", (2, 2), (*sa, *sb)), ("coul", (2, 1), (*sa, *sb)), ("3c2e", (2, 2, 2), (*sa, *sb, *sc))):
    eng.eval_block(key, Ls, *args, R=np.zeros(3))
    t0 = time.perf_counter()
    for _ in range(200):
        eng.eval_block(key, Ls, *args, R=np.zeros(3))
    print(key, "%.1f us per block" % ((time.perf_counter() - t0) / 200 * 1e6))
eng.close()
