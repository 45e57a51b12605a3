"""Memory-safety evidence for the oracle: the golden replays and a crowded random
rollout run against an AddressSanitizer + UBSan build of oracle/*.c (make -C oracle asan) in a child
process; any out-of-bounds access, use-after-free or undefined behaviour aborts the child."""
import os
import subprocess
import sys

import parity

CHILD = r"""
import sys, numpy as np
import parity, pyoracle as po, test_cpu_golden as tg
assert po.build().endswith('liboracle_asan.so')
for path in tg.GOLDEN:
    if not any(k in path for k in ('g_4v4_crowd', 'g_ffa_hoard', 'g_2v2_owned_attack', 'g_ffa_lidar', 'g_1v1_modules')):
        continue
    g = np.load(path); rec = tg.case_config(path)
    seed, env_id, _ = [int(v) for v in g['meta']]
    orc = po.OracleEnv(rec, seed=seed, env_id=env_id)
    for r in range(len(g['kind'])):
        out = orc.reset() if g['kind'][r] == 0 else orc.step(g['actions'][r])
        assert np.array_equal(out['agent'], g['agent'][r])
    orc.close()
rng = np.random.default_rng(0)
rec = parity.make_config('ffa_lidar', auto_reset=True, spawn_grid={'grid_size': 8, 'floor_size': 14}, inventory={'slots': 1})
envs = [po.OracleEnv(rec, seed=1, env_id=e) for e in range(6)]
for o in envs:
    o.reset()
for t in range(300):
    for o in envs:
        a = parity.random_actions(rng, 1, 8)[0]; a[:, 0] = 2
        o.step(a)
st = envs[0].get_state(); envs[0].set_state(st); envs[0].flush_stats()
b = po.OracleBatch(rec, 3, 32, 4); b.reset()
for t in range(40):
    b.step(parity.random_actions(rng, 32, 8))
b.close()
print('SANITIZER-CLEAN')
"""


def test_oracle_under_asan_ubsan():
    import pytest
    root = parity.ROOT
    cc = os.environ.get('ASAN_CC', '/usr/bin/gcc')
    libasan = subprocess.run([cc, '-print-file-name=libasan.so'], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(libasan) or not os.path.exists(libasan):
        pytest.skip('no AddressSanitizer runtime for ' + cc)
    libasan = os.path.realpath(libasan)
    subprocess.check_call(['make', '-C', os.path.join(root, 'oracle'), 'asan', 'ASAN_CC=' + cc], stdout=subprocess.DEVNULL)
    env = dict(os.environ, ORACLE_LIB='liboracle_asan.so', LD_PRELOAD=libasan,
               ASAN_OPTIONS='detect_leaks=0:abort_on_error=1', UBSAN_OPTIONS='halt_on_error=1:print_stacktrace=1',
               PYTHONPATH=os.pathsep.join([os.path.join(root, 'tests'), os.path.join(root, 'oracle'), os.path.join(root, 'gym-ma-survival-2d_b200')]))
    p = subprocess.run([sys.executable, '-c', CHILD], env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and 'SANITIZER-CLEAN' in p.stdout, (p.stdout[-2000:], p.stderr[-4000:])
