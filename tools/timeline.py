"""Overlapped timeline of one learner step (CUDA events around every launch on its own stream; not product code).
    python tools/timeline.py [workload] [option=value ...]      (library options of mal_set_option)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
import ma_league_b200 as M
from ma_league_b200 import _native as nat
from tests.gpu_helpers import seeded_system
from ma_league_b200.synthetic import CONFIGS

wl = sys.argv[1] if len(sys.argv) > 1 else "qmix_5v5_b32"
c = CONFIGS[wl]
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    nat.check(nat.lib().mal_set_option(k.encode(), int(v)), "mal_set_option " + k)
s = seeded_system(c["N"], c["B"], 201, c["mixer"], True, seed=1)
s.learner.use_graphs = False
for i in range(5):
    s.learner.train(s.batch, i, 0)
th.cuda.synchronize()
for rep in range(2):
    nat.profile_begin()
    s.learner.train(s.batch, 0, 0)
    tl = nat.profile_end_timeline()
for name, a, b in tl:
    print("%-34s %8.1f -> %8.1f  (%6.1f us)" % (name, a, b, b - a))
