"""GPU parity tests, part 2 (run with -m gpu on the B200 box): the configurations of BASELINE.json that round 1 left
unpinned — sensitivity levels 1-3, verbosity-3 console text, 5-20 kb pairs at the bridging bands, the calibration pinned
pair by pair, config-5-style synthetic long reads, the free-end-gap entry points (exhaustive / start / end / overlap),
and concurrent per-read ABI calls (request coalescer).  Every expectation is an output of the UNMODIFIED reference
library (tests/golden/make_golden.py), compared byte for byte (field 9, wall-clock ms, masked)."""
import hashlib
import os
import threading

import pytest

from oracle_lib import REF_LIB, load_golden, mask_ms, mask_semi_global

pytestmark = pytest.mark.gpu
SCHEME = (3, -6, -5, -2)


def test_free_end_gap_entry_points_golden(ub):
    """semiGlobalAlignmentExhaustive / startAlignment / endAlignment / overlapAlignment."""
    d = load_golden('freeend.json.gz')
    bad = []
    for k, c in enumerate(d['cases']):
        sc = tuple(c['scheme'])
        if c['kind'] == 'exhaustive':
            got = mask_ms(ub.semi_global_alignment_exhaustive(c['s1'], c['s2'], sc))
        elif c['kind'] == 'start':
            got = ub.start_seq_alignment(c['s1'], c['s2'], sc)
        elif c['kind'] == 'end':
            got = ub.end_seq_alignment(c['s1'], c['s2'], sc)
        else:
            got = '%d,%d' % ub.overlap_alignment(c['s1'], c['s2'], sc, c['guess'])
        if got != c['result']:
            bad.append((k, c['kind'], sc, len(c['s1']), len(c['s2']), str(got)[:80], str(c['result'])[:80]))
    assert len(d['cases']) >= 160
    assert not bad, bad[:5]


def test_pairs_at_bridging_sizes_golden(ub):
    """5 / 10 / 20 kb pairs at band 1000 / 500 / 50 (path_finding.py:71,323,334, bridge_long_read_simple.py:485) and
    an unbanded 5 kb pair, both entry points, single-call ABI and batch ABI."""
    d = load_golden('pairs_large.json.gz')
    groups = {}
    for c in d['cases']:
        groups.setdefault((tuple(c['scheme']), c['banded'], c['band']), []).append(c)
    n = 0
    for (sc, banded, band), cases in groups.items():
        g = ub.fully_global_alignment_batch([c['s1'] for c in cases], [c['s2'] for c in cases], sc, banded, band)
        p = ub.path_alignment_batch([c['s1'] for c in cases], [c['s2'] for c in cases], sc, banded, band)
        for c, gg, pp in zip(cases, g, p):
            assert mask_ms(gg) == c['global'], ('global', banded, band, len(c['s1']), len(c['s2']))
            assert mask_ms(pp) == c['path'], ('path', banded, band, len(c['s1']), len(c['s2']))
            n += 1
    c = max(d['cases'], key=lambda c: len(c['s1']) * (c['band'] if c['banded'] else 10 ** 9))
    assert mask_ms(ub.fully_global_alignment(c['s1'], c['s2'], tuple(c['scheme']), c['banded'], c['band'])) == c['global']
    assert mask_ms(ub.path_alignment(c['s1'], c['s2'], tuple(c['scheme']), c['banded'], c['band'])) == c['path']
    assert n >= 19 and max(len(c['s1']) for c in d['cases']) >= 20000


def test_calibration_pinned_pair_by_pair(ub, monkeypatch):
    """getRandomSequenceAlignmentScores with a fixed generator seed: every pair it aligns gives the reference's result
    string, and the returned "mean,sd" equals getMeanAndStDev over the reference's scaled scores to the last digit."""
    d = load_golden('calibration_pairs.json.gz')
    for s in d['sets']:
        sc = tuple(s['scheme'])
        a, b = ub.calibration_pairs(s['length'], s['n'], s['seed'])
        assert [a[0], b[0]] == s['first_pair']
        got = ub.fully_global_alignment_batch(a, b, sc, False, 0)
        assert [mask_ms(g) for g in got] == s['results'], (s['length'], s['n'])
        monkeypatch.setenv('UNICYCLER_B200_SEED', str(s['seed']))
        m, mm, go, ge = sc
        ptr = ub.load_library().getRandomSequenceAlignmentScores(s['length'], s['n'], m, mm, go, ge)
        from unicycler_b200.wrappers import _to_str
        assert _to_str(ptr) == s['mean_sd'], (s['length'], s['n'])


def test_synthetic_long_reads_config5_golden(ub):
    """BASELINE configs[4] inputs at reduced depth: 16 reads x 20 kb (15 % errors) vs a random 1 Mbp reference,
    regenerated from the seed and checked by checksum; expected strings from the reference library."""
    from make_golden import synth5_inputs
    d = load_golden('semiglobal_synth5.json.gz')
    ref, reads = synth5_inputs(d['ref_len'], d['n_reads'], d['read_len'], d['seed'])
    assert hashlib.sha1(ref.encode()).hexdigest() == d['ref_sha1']
    assert hashlib.sha1(''.join(r[1] for r in reads).encode()).hexdigest() == d['reads_sha1']
    h = ub.new_ref_seqs()
    ub.add_ref_seq(h, 'ref', ref)
    out = ub.semi_global_alignment_batch([r[0] for r in reads], [r[1] for r in reads], [r[2] for r in reads], h,
                                         tuple(d['scheme']), d['sensitivity'])
    ub.delete_ref_seqs(h)
    bad = [r[0] for r, o in zip(reads, out) if mask_semi_global(o) != d['expected'][r[0]]]
    assert not bad, bad
    assert sum(len(o.split(';')) - 1 for o in out) >= len(reads)


@pytest.mark.parametrize('setname', ['small', 'sample', 'contained'])
def test_sensitivity_levels_and_console_text_golden(ub, setname):
    """Sensitivity 1-3 (band 50 / 75 / 100, k 10 / 9 / 8, up to 16 line traces: settings.h:17-42,
    semi_global_align.cpp:163-192) through the batch ABI, and verbosity 3 (console text after the last ';',
    semi_global_align.cpp:36-63,144-152) through the per-read ABI."""
    g = load_golden('semiglobal_sensitivity.json.gz')
    d = load_golden('semiglobal_%s.json.gz' % setname)
    reads = {r[0]: r for r in d['reads']}
    h = ub.new_ref_seqs()
    for name, seq in d['refs']:
        ub.add_ref_seq(h, name, seq)
    entries = g['sets'][setname]
    sc = tuple(g['scheme'])
    bad = []
    for level in (1, 2, 3):
        es = [e for e in entries if e['verbosity'] == 0 and e['sensitivity'] == level]
        assert es
        rs = [reads[e['read']] for e in es]
        out = ub.semi_global_alignment_batch([r[0] for r in rs], [r[1] for r in rs], [r[2] for r in rs], h, sc, level)
        bad += [(e['read'], level) for e, o in zip(es, out) if mask_semi_global(o) != e['expected']]
    m, mm, go, ge = sc
    for e in entries:
        if e['verbosity'] == 0:
            continue
        r = reads[e['read']]
        o = ub.semi_global_alignment(r[0], r[1], e['verbosity'], r[2], h, m, mm, go, ge, 0.0, False, e['sensitivity'])
        assert o.split(';')[-1] != ''   # there is console text
        if mask_semi_global(o) != e['expected']:
            bad.append((e['read'], e['sensitivity'], 'verbosity %d' % e['verbosity']))
    ub.delete_ref_seqs(h)
    assert not bad, bad


def test_concurrent_per_read_calls_are_coalesced(ub):
    """Eight threads call the per-read ABI at once, as unicycler_align.py:203-225 does; every call returns the
    reference's string and the coalescer merged requests into shared device batches."""
    d = load_golden('semiglobal_sample.json.gz')
    h = ub.new_ref_seqs()
    for name, seq in d['refs']:
        ub.add_ref_seq(h, name, seq)
    reads = [r for r in d['reads'] if r[0] in d['expected']]
    m, mm, go, ge = d['scheme']
    out = [None] * len(reads)
    nxt = [0]
    lock = threading.Lock()
    before = ub.coalescer_stats()

    def worker():
        while True:
            with lock:
                k = nxt[0]
                nxt[0] += 1
            if k >= len(reads):
                return
            r = reads[k]
            out[k] = ub.semi_global_alignment(r[0], r[1], 0, r[2], h, m, mm, go, ge, 0.0, False, 0)

    threads = [threading.Thread(target=worker) for _ in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    after = ub.coalescer_stats()
    ub.delete_ref_seqs(h)
    bad = [r[0] for r, o in zip(reads, out) if mask_semi_global(o) != d['expected'][r[0]]]
    assert not bad, bad
    req, bat = after['requests'] - before['requests'], after['batches'] - before['batches']
    assert req >= len(reads) - 2 and bat >= 1
    assert bat < req, (bat, req)   # some calls shared a launch


@pytest.mark.skipif(not os.path.isfile(REF_LIB), reason='oracle/_ref not built')
def test_free_end_gap_entry_points_vs_reference_library(ub):
    """Fresh random cases against the reference library itself (not only the committed vectors)."""
    import random
    from refdriver import AbiLib
    ref = AbiLib(REF_LIB)
    rng = random.Random(31)

    def rs(n):
        return ''.join(rng.choice('ACGT') for _ in range(n))

    for it in range(60):
        a = rs(rng.randint(5, 600))
        b = rs(rng.randint(0, 200)) + a[rng.randint(0, len(a) // 2):] + rs(rng.randint(0, 200))
        if it % 3 == 0:
            b = rs(rng.randint(5, 900))
        assert mask_ms(ub.semi_global_alignment_exhaustive(a, b, SCHEME)) == mask_ms(ref.exhaustive(a, b, SCHEME))
        assert ub.start_seq_alignment(a, b, SCHEME) == ref.start(a, b, SCHEME)
        assert ub.end_seq_alignment(a, b, SCHEME) == ref.end(a, b, SCHEME)
        g = rng.randint(0, 300)
        assert '%d,%d' % ub.overlap_alignment(a, b, SCHEME, g) == ref.overlap(a, b, SCHEME, g)


@pytest.mark.skipif(not os.path.isfile(REF_LIB), reason='oracle/_ref not built')
def test_unbanded_pair_with_more_than_2_31_cells_vs_reference_library(ub):
    """The calibration sweep of BASELINE configs[3] goes to 50 kb: (47 001)^2 cells do not fit 32-bit matrix positions
    (the reference uses size_t).  One such pair of the calibration's own generator against the reference library
    (~35 s on one host core)."""
    from refdriver import AbiLib
    ref = AbiLib(REF_LIB)
    a, b = ub.calibration_pairs(47000, 1, 11)
    got = ub.fully_global_alignment(a[0], b[0], SCHEME, False, 0)
    want = ref.fully_global(a[0], b[0], SCHEME, False, 0)
    assert mask_ms(got) == mask_ms(want)


@pytest.mark.skipif(not os.path.isfile(REF_LIB), reason='oracle/_ref not built')
def test_gap_area_guard_vs_reference_library(ub):
    """Two matching blocks separated by unrelated sequence: when the largest gap between chained seeds exceeds 1e8 cells
    the reference gives up on the range (getMaxSeedChainGapArea, semi_global_align.cpp:286-291,321-347); just below and
    above that size it aligns.  Same outcome, same strings."""
    import random
    from refdriver import AbiLib
    ref_lib = AbiLib(REF_LIB)
    rng = random.Random(5)

    def rs(n):
        return ''.join(rng.choice('ACGT') for _ in range(n))

    outcomes = []
    for gap in (9000, 10500, 12000, 15000):
        a, b = rs(1500), rs(1500)
        read = a + rs(gap) + b
        ref = rs(300) + a + rs(gap) + b + rs(300)
        hits = '0,%d,+,ref,300,%d' % (len(read), 300 + len(read))
        h = ub.new_ref_seqs()
        ub.add_ref_seq(h, 'ref', ref)
        got = ub.semi_global_alignment('r', read, 0, hits, h, 3, -6, -5, -2, 0.0, False, 0)
        ub.delete_ref_seqs(h)
        hr = ref_lib.new_refs([('ref', ref)])
        want = ref_lib.semi_global('r', read, hits, hr, SCHEME)
        ref_lib.delete_refs(hr)
        assert mask_semi_global(got) == mask_semi_global(want), gap
        outcomes.append(len(want.split(';')) - 1)
    assert 0 in outcomes and max(outcomes) >= 1   # the guard fired for some sizes and not for others


def test_output_stream_overflow_is_rerun_with_a_larger_stream(tmp_path):
    """A job whose segment stream overflows (many tied tracebacks in the reference's terms) reports JOB_OUT_OVERFLOW
    and the engine reruns it with an eight times larger stream (Engine::end).  Forced here by starting with a thousandth
    of the planned capacity (test hook UNICYCLER_B200_OUT_CAP_DIV, read once per process: subprocess)."""
    import subprocess
    import sys
    from oracle_lib import ROOT
    code = '''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import unicycler_b200 as ub
from oracle_lib import golden_chain_jobs, load_golden, mask_ms
d = load_golden("semiglobal_contained.json.gz")
jobs = golden_chain_jobs(d)
got = ub.chain_alignment_batch(jobs, tuple(d["scheme"]), jobs[0]["band"])
bad = [j["readName"] for j, g in zip(jobs, got) if mask_ms(g) != j["result"]]
assert not bad, bad
st = ub.last_stats()
assert st["launches"] >= 2, st      # the first launch overflowed, the rerun produced the results
''' % (ROOT, os.path.join(ROOT, 'tests'))
    env = dict(os.environ, UNICYCLER_B200_OUT_CAP_DIV='1000')
    r = subprocess.run([sys.executable, '-c', code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert r.returncode == 0, r.stdout.decode()[-2000:]
