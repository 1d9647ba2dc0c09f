"""Diagnostic (not product code): per-tensor gradient errors of the learner vs the fp64 oracle at a given shape under
different library switches.   python tools/diag_parity.py N B TT [mixer]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch as th
from oracle import np_oracle as O
from tests.gpu_helpers import seeded_system, np_params, np_batch, split_grad
from tests.helpers import rel_err, max_err
from ma_league_b200 import _native as nat

N, B, TT = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mixer = sys.argv[4] if len(sys.argv) > 4 else "qmix"
s = seeded_system(N, B, TT, mixer, True, seed=43)
L = s.learner
t0 = time.time()
ref = O.learner_forward_backward(np_params(s.mac.agent), np_params(L.target_mac.agent),
                                 np_params(L.mixer) if mixer == "qmix" else None,
                                 np_params(L.target_mixer) if mixer == "qmix" else None, np_batch(s.batch),
                                 mixer=mixer, double_q=True, gamma=s.args.gamma, dtype=np.float64)
print("oracle %.1fs" % (time.time() - t0), flush=True)
refg = {("agent." + k): v for k, v in ref["agent_grads"].items()}
refg.update({("mixer." + k): v for k, v in ref["mixer_grads"].items()})
variants = [[], [("reduce_tc", 0)], [("overlap", 0)], [("pdl", 0)], [("tc_pipelined", 0)], [("fuse_agent_in", 0)],
            [("reduce_tc", 0), ("tc_pipelined", 0)], [("tensor_cores", 0)]]
lib = nat.lib()
for var in variants:
    for k, v in var:
        lib.mal_set_option(k.encode(), v)
    g = split_grad(L.forward_backward(s.batch), L)
    th.cuda.synchronize()
    it = {k: v.cpu().numpy() for k, v in L.intermediates(s.batch).items()}
    flips = int((it["argmax"].astype(np.int64) != ref["argmax"]).sum())
    bad = {k: (rel_err(g[k], refg[k]), max_err(g[k], refg[k])) for k in refg}
    worst = sorted(bad.items(), key=lambda kv: -kv[1][0])[:6]
    print(var or "default", "flips", flips, "q_tot %.1e" % rel_err(it["q_tot"], ref["q_tot"]),
          " ".join("%s=%.1e/%.1e" % (k, a, b) for k, (a, b) in worst), flush=True)
    for k, v in var:
        lib.mal_set_option(k.encode(), 1)
