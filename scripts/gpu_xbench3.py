import sys, os, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
from argsim_b200 import _lib
out = {}
for method in (0, 3, 4, 5):
    for groups in (1, 6):
        for rows in (1, 4, 8, 16, 24, 32):
            try:
                cyc, mc = _lib.bench_exchange(method, groups, rows, 3000)
                out['m%d_g%d_r%d' % (method, groups, rows)] = cyc
                print('method', method, 'groups', groups, 'rows', rows, 'cycles/round %.0f' % cyc, 'max_clusters', mc, flush=True)
            except Exception as e:
                print('method', method, 'groups', groups, 'rows', rows, 'FAILED', str(e)[:200], flush=True)
json.dump(out, open(os.path.join(R, 'gpurun_out', 'xbench3.json'), 'w'), indent=1)
