import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
from argsim_b200 import _lib
for rep in range(2):
    for method in (32, 96):
        for groups in (1, 4, 8):
            for rows in (1, 4):
                cyc, _ = _lib.bench_exchange(method, groups, rows, 4000)
                print('method', method, 'groups', groups, 'rows', rows, 'cycles/round %.0f' % cyc, flush=True)
