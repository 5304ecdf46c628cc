"""Developer probe: device-resident kernel time of the heavy / light part of the sample_data chain jobs (reads sorted by
length, cut at a fraction of the total length), to be run with UNICYCLER_B200_CTAS=74 and without."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import unicycler_b200 as ub
from oracle_lib import golden_chain_jobs, load_golden
frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
d = load_golden('semiglobal_sample.json.gz')
jobs = golden_chain_jobs(d)
by_read = {}
for j in jobs:
    by_read.setdefault(j['readName'][:-1], []).append(j)
reads = sorted(by_read, key=lambda r: -len(by_read[r][0]['readSeq']))
total = sum(len(by_read[r][0]['readSeq']) for r in reads)
acc, heavy, light = 0, [], []
for r in reads:
    (heavy if acc < frac * total else light).append(r)
    acc += len(by_read[r][0]['readSeq'])
for tag, part in (('heavy', heavy), ('light', light), ('all', reads)):
    js = [j for r in part for j in by_read[r]]
    if not js:
        continue
    cells = sum(ub.chain_cells(len(j['readSeq']), len(j['refSeq']), j['seeds'], j['band'])[0] for j in js)
    b = ub.ChainBench(js, tuple(d['scheme']), js[0]['band'])
    b.run_steps(2)
    ms = b.run_steps(4) / 4
    b.finish(False)
    print('HALF ctas=%s frac=%.2f %s: reads=%d jobs=%d cells=%.3g kernel_ms=%.2f' % (os.environ.get('UNICYCLER_B200_CTAS', 'all'), frac, tag, len(part), len(js), cells, ms), flush=True)
