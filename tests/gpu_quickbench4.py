import sys
import gpu_quickbench as q
v = sys.argv[1] if len(sys.argv) > 1 else '2v2'
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
for r in range(2):
    q.run(v, N, steps=600, warm=100, prof=False)
