"""Host-side cost of QLearner.train(): cProfile over replayed-graph steps (the GPU is not the bottleneck here)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from tests.gpu_helpers import seeded_system

s = seeded_system(5, 32, 201, "qmix", True, seed=0)
L = s.learner
for i in range(5):
    L.train(s.batch, i, 0)
th.cuda.synchronize()
t0 = time.perf_counter()
for i in range(300):
    L.train(s.batch, i, 0)
t1 = time.perf_counter()
th.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue per train(): %.1f us   (device-bound total %.1f us/step)" % ((t1 - t0) / 300 * 1e6, (t2 - t0) / 300 * 1e6))
pr = cProfile.Profile()
pr.enable()
for i in range(300):
    L.train(s.batch, i, 0)
pr.disable()
th.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
