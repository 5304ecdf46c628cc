import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import unicycler_b200 as ub
from unicycler_b200 import wrappers as W
from oracle_lib import load_golden
d = load_golden('semiglobal_sample.json.gz')
h = ub.new_ref_seqs()
for n_, s_ in d['refs']:
    ub.add_ref_seq(h, n_, s_)
reads = [r for r in d['reads'] if r[0] in d['expected']]
names, seqs, hits = [r[0] for r in reads], [r[1] for r in reads], [r[2] for r in reads]
L = W.load_library()
for rep in range(8):
    t0 = time.perf_counter()
    a, b, c = W._cstrs(names), W._cstrs(seqs), W._cstrs(hits)
    out = (ctypes.c_void_p * len(names))()
    t1 = time.perf_counter()
    L.ub200_semiGlobalAlignmentBatch(len(names), a, b, c, h, 3, -6, -5, -2, 0, out)
    t2 = time.perf_counter()
    res = [W._to_str(p) for p in out]
    t3 = time.perf_counter()
    print('marshal %.2f call %.2f results %.2f ms (%d bytes)' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, sum(len(x) for x in res)))
