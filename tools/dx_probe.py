import sys, os
sys.path.insert(0, os.getcwd())
sys.argv = ["x"]
import importlib.util
spec = importlib.util.spec_from_file_location("gb", "tools/gemm_bench.py")
gb = importlib.util.module_from_spec(spec)
src = open("tools/gemm_bench.py").read().split('if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "light":')[0]
exec(compile(src, "gb", "exec"), gb.__dict__)
M = 4116480 // 4
for epi in (2, 0):
    for dbg in (0, 1, 2, 8, 3):
        gb.lib.mal_set_option(b"tc_dbg", dbg)
        print("epi", epi, "dbg", dbg, end="  ")
        gb.bench(M, 192, 64, epi=epi, w_trans=0, reps=5)
gb.lib.mal_set_option(b"tc_dbg", 0)
gb.bench(M, 64, 64, epi=2, reps=5)
gb.bench(M, 64, 64, epi=0, reps=5)
gb.bench(M, 320, 64, epi=1, reps=5)
