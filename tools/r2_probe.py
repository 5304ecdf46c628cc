"""Developer probe: device-resident kernel time of the sample_data chain jobs replicated k times (k = 1, 2, 4, 8...)
— the same rectangles in a throughput-sized batch — plus the kernel's own timeline for k = 1."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import unicycler_b200 as ub
from oracle_lib import golden_chain_jobs, load_golden, mask_ms

name = sys.argv[1] if len(sys.argv) > 1 else 'sample'
reps = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else '1,2,4,8').split(',')]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
d = load_golden('semiglobal_%s.json.gz' % name)
jobs = golden_chain_jobs(d)
cells1 = sum(ub.chain_cells(len(j['readSeq']), len(j['refSeq']), j['seeds'], j['band'])[0] for j in jobs)
for k in reps:
    b = ub.ChainBench(jobs * k, tuple(d['scheme']), jobs[0]['band'])
    b.run_steps(2)
    ms = b.run_steps(steps) / steps
    res = b.finish(True)
    bad = 0
    for j, g in zip(jobs * k, res):
        f = j['result'].split(',', 9)
        if len(f) >= 10:
            f[0], f[1] = 'ref', '+'
            f[4], f[5] = str(int(f[4]) - j['refOffset']), str(int(f[5]) - j['refOffset'])
        if mask_ms(g) != ','.join(f):
            bad += 1
    print('PROBE set=%s x%d jobs=%d cells=%.4g kernel_ms=%.3f GCUPS=%.1f int_frac(17 ops @37.09e12)=%.3f bad=%d' %
          (name, k, len(jobs) * k, cells1 * k, ms, cells1 * k / ms / 1e6, cells1 * k * 17 / (ms * 1e-3) / 37.09e12, bad),
          flush=True)
