import sys
import gpu_quickbench as q
for rep in range(3):
    for v, N in (('2v2', 16384), ('2v2', 32768)):
        q.run(v, N, prof=(rep == 2))
