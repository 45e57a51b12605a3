import sys
import gpu_quickbench as q
v = sys.argv[1]; N = int(sys.argv[2])
q.run(v, N, steps=300, warm=100, prof=False)
