"""Developer aid: runs the parity categories on a GPU and logs mismatch counts/details (no assert-stop)."""
import os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import unicycler_b200 as ub
from oracle_lib import Oracle, golden_chain_jobs, load_golden, mask_ms, mask_semi_global
orc = Oracle()
what = sys.argv[1:] or ['global', 'small', 'contained', 'tough']
t0 = time.time()
if 'global' in what:
    d = load_golden('global_path.json.gz')
    groups = {}
    for c in d['cases']:
        if len(c['s1']) < 2 or len(c['s2']) < 2: continue
        groups.setdefault((tuple(c['scheme']), c['banded'], c['band']), []).append(c)
    badg = badp = n = 0
    for (sc, banded, band), cases in groups.items():
        g = ub.fully_global_alignment_batch([c['s1'] for c in cases], [c['s2'] for c in cases], sc, banded, band)
        p = ub.path_alignment_batch([c['s1'] for c in cases], [c['s2'] for c in cases], sc, banded, band)
        for c, gg, pp in zip(cases, g, p):
            n += 1
            if mask_ms(gg) != c['global']:
                badg += 1
                if badg <= 4: print('GLOBAL MISMATCH', sc, banded, band, len(c['s1']), len(c['s2']), '\n  got ', mask_ms(gg)[:160], '\n  want', c['global'][:160])
            if mask_ms(pp) != c['path']:
                badp += 1
                if badp <= 4: print('PATH MISMATCH', sc, banded, band, len(c['s1']), len(c['s2']), '\n  got ', mask_ms(pp)[:160], '\n  want', c['path'][:160])
    print('global/path cases', n, 'bad global', badg, 'bad path', badp, 'time %.1f' % (time.time() - t0), ub.last_stats(), flush=True)
for setname in ['small', 'contained', 'tough', 'sample']:
    if setname not in what: continue
    t0 = time.time()
    d = load_golden('semiglobal_%s.json.gz' % setname)
    jobs = golden_chain_jobs(d)
    got = ub.chain_alignment_batch(jobs, tuple(d['scheme']), jobs[0]['band'])
    st = ub.last_stats()
    bad = 0
    for k, (j, g) in enumerate(zip(jobs, got)):
        if mask_ms(g) != j['result']:
            bad += 1
            if bad <= 4: print('CHAIN MISMATCH', setname, k, j['readName'], j['refName'], len(j['readSeq']), len(j['refSeq']), len(j['seeds']), '\n  got ', mask_ms(g)[:200], '\n  want', j['result'][:200])
    print('chain', setname, 'jobs', len(jobs), 'bad', bad, 'time %.2f' % (time.time() - t0), st, 'GCUPS(kernel) %.2f' % (st['cells'] / max(st['kernel_ms'], 1e-9) / 1e6), flush=True)
    # end-to-end through semiGlobalAlignment
    t0 = time.time()
    h = ub.new_ref_seqs()
    for name, seq in d['refs']: ub.add_ref_seq(h, name, seq)
    reads = [r for r in d['reads'] if r[0] in d['expected']]
    out = ub.semi_global_alignment_batch([r[0] for r in reads], [r[1] for r in reads], [r[2] for r in reads], h, tuple(d['scheme']), d['sensitivity'])
    bad = sum(1 for r, o in zip(reads, out) if mask_semi_global(o) != d['expected'][r[0]])
    print('semi-global e2e', setname, 'reads', len(reads), 'bad', bad, 'time %.2f' % (time.time() - t0), ub.last_stats(), flush=True)
