"""Per-kernel times of one learner step, kernels timed alone (not product code).  python tools/kernels.py WORKLOAD [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from ma_league_b200 import _native as nat
from tests.gpu_helpers import seeded_system
from ma_league_b200.synthetic import CONFIGS
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
wl = sys.argv[1]
c = CONFIGS[wl]
B = int(sys.argv[2]) if len(sys.argv) > 2 else c["B"]
s = seeded_system(c["N"], B, 201, c["mixer"], True, seed=1)
s.learner.use_graphs = False
for i in range(3):
    s.learner.train(s.batch, i, 0)
th.cuda.synchronize()
e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
e0.record()
for i in range(5):
    s.learner.train(s.batch, i, 0)
e1.record(); th.cuda.synchronize()
print("%s B=%d: %.3f ms/step (eager, overlapped)" % (wl, B, e0.elapsed_time(e1) / 5))
nat.lib().mal_set_option(b"overlap", 0)
nat.profile_begin()
for i in range(3):
    s.learner.train(s.batch, i, 0)
prof = nat.profile_end()
d = bench.workload_dims(wl)
kb = bench.kernel_bytes(d, c["mixer"], B)
peak, _ = bench.hbm_peak()
tot = sum(v[1] for v in prof.values())
for k, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    per = ms / n
    print("  %-32s %9.1f us  share %.3f  %s" % (k, per * 1e3, ms / tot, ("%.0f GB/s = %.3f of HBM peak" % (kb[k] / (per * 1e-3) / 1e9, kb[k] / (per * 1e-3) / 1e9 / peak)) if k in kb else ""))
