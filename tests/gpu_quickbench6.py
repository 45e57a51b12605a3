import gpu_quickbench as q
for v, N in (('2v2', 16384), ('1v1_heal_only', 4096), ('1v1', 16384), ('ffa', 32768), ('ffa_lidar', 32768)):
    q.run(v, N, steps=500, warm=100, prof=False)
