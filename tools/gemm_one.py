import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.gemm_bench import bench
bench(int(sys.argv[1]) if len(sys.argv) > 1 else 257280, 64, 192, reps=2)
