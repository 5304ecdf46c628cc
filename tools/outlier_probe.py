"""Developer probe: many batch calls on sample_data with the per-stage profile; prints the calls that took long."""
import os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import unicycler_b200 as ub
    from oracle_lib import load_golden
    d = load_golden('semiglobal_sample.json.gz')
    h = ub.new_ref_seqs()
    for n_, s_ in d['refs']:
        ub.add_ref_seq(h, n_, s_)
    reads = [r for r in d['reads'] if r[0] in d['expected']]
    args = ([r[0] for r in reads], [r[1] for r in reads], [r[2] for r in reads], h, tuple(d['scheme']), 0)
    for rep in range(int(sys.argv[2])):
        t0 = time.perf_counter()
        ub.semi_global_alignment_batch(*args)
        sys.stderr.write('[call] %d %.1f ms\n' % (rep, (time.perf_counter() - t0) * 1e3))
else:
    n = sys.argv[1] if len(sys.argv) > 1 else '80'
    env = dict(os.environ, UNICYCLER_B200_PROFILE='1')
    out = subprocess.run([sys.executable, __file__, 'child', n], env=env, stderr=subprocess.PIPE).stderr.decode().split('\n')
    calls, block = [], []
    for line in out:
        block.append(line)
        if line.startswith('[call]'):
            calls.append((float(line.split()[2]), block))
            block = []
    ts = sorted(c[0] for c in calls[3:])
    print('calls %d: min %.1f median %.1f p90 %.1f max %.1f' % (len(ts), ts[0], ts[len(ts) // 2], ts[int(len(ts) * 0.9)], ts[-1]))
    for t, blk in calls[3:]:
        if t > 1.3 * ts[len(ts) // 2]:
            print('--- slow call %.1f ms' % t)
            for l in blk:
                if l.startswith('[ub200 upload] staging') or l.startswith('[ub200 engine]') or l.startswith('[ub200 host] reads') or l.startswith('[ub200 timeline] last spine') or l.startswith('[ub200 fetch]'):
                    print('   ', l[:230])
