import sys
import gpu_quickbench as q
v = sys.argv[1] if len(sys.argv) > 1 else '2v2'
for N in (1024, 2048, 4096, 8192, 16384, 32768):
    q.run(v, N, steps=500, warm=100, prof=False)
