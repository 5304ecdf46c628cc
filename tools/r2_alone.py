"""Developer probe: kernel time of the heaviest sample_data chain jobs when each runs ALONE (its latency floor)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import unicycler_b200 as ub
from oracle_lib import golden_chain_jobs, load_golden
d = load_golden('semiglobal_sample.json.gz')
jobs = golden_chain_jobs(d)
cells = [ub.chain_cells(len(j['readSeq']), len(j['refSeq']), j['seeds'], j['band'])[0] for j in jobs]
order = sorted(range(len(jobs)), key=lambda k: -cells[k])
which = [int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else order[:12]
for k in which:
    b = ub.ChainBench([jobs[k]], tuple(d['scheme']), jobs[k]['band'])
    b.run_steps(2)
    ms = b.run_steps(3) / 3
    plan = ub.chain_plan(len(jobs[k]['readSeq']), len(jobs[k]['refSeq']), jobs[k]['seeds'], jobs[k]['band'])
    big = sorted(((p[1], p[2]) for p in plan if p[2] > 256 or p[1] * p[2] > 40000), key=lambda x: -x[0] * x[1])[:4]
    print('ALONE job %d cells %.3g grids %d: %.3f ms  biggest grids %s' % (k, cells[k], len(plan), ms, big), flush=True)
