"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against
(1) the golden vectors of the unmodified reference, (2) the oracle on seeded random inputs, and
(3) size-independent properties at full size."""
import os
import random

import pytest

from oracle_lib import REF_LIB, golden_chain_jobs, load_golden, mask_ms, mask_semi_global

pytestmark = pytest.mark.gpu
SCHEME = (3, -6, -5, -2)


def _mutate(s, rate, rng):
    out = []
    for c in s:
        r = rng.random()
        if r < rate / 3:
            out.append(rng.choice('ACGT'))
        elif r < 2 * rate / 3:
            continue
        elif r < rate:
            out.append(c)
            out.append(rng.choice('ACGT'))
        else:
            out.append(c)
    return ''.join(out) or 'AC'


def test_global_path_golden(ub):
    d = load_golden('global_path.json.gz')
    groups = {}
    for c in d['cases']:
        if len(c['s1']) < 2 or len(c['s2']) < 2:
            continue
        groups.setdefault((tuple(c['scheme']), c['banded'], c['band']), []).append(c)
    n = 0
    for (sc, banded, band), cases in groups.items():
        g = ub.fully_global_alignment_batch([c['s1'] for c in cases], [c['s2'] for c in cases], sc, banded, band)
        p = ub.path_alignment_batch([c['s1'] for c in cases], [c['s2'] for c in cases], sc, banded, band)
        for c, gg, pp in zip(cases, g, p):
            assert mask_ms(gg) == c['global'], ('global', sc, banded, band, c['s1'][:40], c['s2'][:40])
            assert mask_ms(pp) == c['path'], ('path', sc, banded, band, c['s1'][:40], c['s2'][:40])
            n += 1
    assert n > 600


def test_single_call_abi_matches_batch(ub):
    s1, s2 = 'ACGTACGTTTGACCAGTAGGATTACA', 'ACGTACGTTGACCAGTTAGGATACA'
    assert mask_ms(ub.fully_global_alignment(s1, s2, SCHEME, False, 0)) == \
        mask_ms(ub.fully_global_alignment_batch([s1], [s2], SCHEME, False, 0)[0])
    assert mask_ms(ub.path_alignment(s1, s2, SCHEME, True, 10)) == \
        mask_ms(ub.path_alignment_batch([s1], [s2], SCHEME, True, 10)[0])


def test_global_path_random_vs_oracle(ub, oracle):
    rng = random.Random(99)
    schemes = [(3, -6, -5, -2), (1, -1, -1, -1), (5, -4, -8, -6), (1, -3, -5, -2), (3, -6, -2, -5)]
    for sc in schemes:
        for banded, band in ((False, 0), (True, 7), (True, 50), (True, 600)):
            s1s, s2s = [], []
            for _ in range(40):
                L = rng.choice([3, 17, 64, 130, 513, 700, 1100, 2300])
                a = ''.join(rng.choice('ACGT') for _ in range(L))
                b = _mutate(a, rng.choice([0.02, 0.15, 0.3]), rng)
                if rng.random() < 0.3:
                    b += ''.join(rng.choice('ACGT') for _ in range(rng.randint(1, 300)))
                if rng.random() < 0.3:
                    a, b = b, a
                s1s.append(a)
                s2s.append(b)
            g = ub.fully_global_alignment_batch(s1s, s2s, sc, banded, band)
            p = ub.path_alignment_batch(s1s, s2s, sc, banded, band)
            for a, b, gg, pp in zip(s1s, s2s, g, p):
                assert mask_ms(gg) == mask_ms(oracle.fully_global(a, b, sc, banded, band)), (sc, banded, band, len(a), len(b))
                assert mask_ms(pp) == mask_ms(oracle.path(a, b, sc, banded, band)), (sc, banded, band, len(a), len(b))


@pytest.mark.parametrize('setname', ['small', 'contained', 'tough'])
def test_chain_golden(ub, setname):
    d = load_golden('semiglobal_%s.json.gz' % setname)
    jobs = golden_chain_jobs(d)
    got = ub.chain_alignment_batch(jobs, tuple(d['scheme']), jobs[0]['band'])
    bad = [(j['readName'], j['refName']) for j, g in zip(jobs, got) if mask_ms(g) != j['result']]
    assert not bad, bad[:5]


@pytest.mark.parametrize('setname', ['small', 'contained', 'tough', 'sample'])
def test_semi_global_end_to_end_golden(ub, setname):
    """semiGlobalAlignment through the reference ABI (seeding on the host, DP on the GPU) against the
    unmodified reference's output strings."""
    d = load_golden('semiglobal_%s.json.gz' % setname)
    h = ub.new_ref_seqs()
    for name, seq in d['refs']:
        ub.add_ref_seq(h, name, seq)
    reads = [r for r in d['reads'] if r[0] in d['expected']]
    out = ub.semi_global_alignment_batch([r[0] for r in reads], [r[1] for r in reads], [r[2] for r in reads], h,
                                         tuple(d['scheme']), d['sensitivity'])
    bad = [r[0] for r, o in zip(reads, out) if mask_semi_global(o) != d['expected'][r[0]]]
    # the per-read ABI call must give the same string as the batch call
    r0 = reads[0]
    m, mm, go, ge = d['scheme']
    single = ub.semi_global_alignment(r0[0], r0[1], 0, r0[2], h, m, mm, go, ge, 0.0, False, d['sensitivity'])
    ub.delete_refs = None
    assert mask_semi_global(single) == d['expected'][r0[0]]
    assert not bad, bad


def test_chain_linear_gaps_and_other_schemes_vs_oracle(ub, oracle):
    """Schemes the golden files do not cover (linear gap model, different penalties) on real seed chains."""
    d = load_golden('semiglobal_contained.json.gz')
    jobs = golden_chain_jobs(d)[:3]
    for sc in [(1, -1, -1, -1), (5, -4, -8, -6), (2, -3, -3, -3)]:
        got = ub.chain_alignment_batch(jobs, sc, 25)
        for j, g in zip(jobs, got):
            want = oracle.chain(j['readSeq'], j['refSeq'], j['seeds'], sc, 25, j['readName'], j['refName'], j['refOffset'])
            assert mask_ms(g) == mask_ms(want), (sc, j['readName'])


def test_calibration_statistics(ub):
    """getRandomSequenceAlignmentScores: the reference seeds from std::random_device, so parity is
    statistical: (mean, sd) for 3,-6,-5,-2 must match the precomputed table (unicycler_align.py:497-498:
    61.656918, 1.314624) within sampling error; identical pairs are covered by the global tests."""
    os.environ['UNICYCLER_B200_SEED'] = '42'
    mean, sd = ub.get_random_sequence_alignment_mean_and_std_dev(100, 25000, SCHEME)
    assert abs(mean - 61.656918) < 0.05
    assert abs(sd - 1.314624) < 0.05
    mean2, sd2 = ub.get_random_sequence_alignment_mean_and_std_dev(100, 25000, SCHEME)
    assert (mean, sd) == (mean2, sd2)  # deterministic under UNICYCLER_B200_SEED


def test_full_size_properties(ub):
    """BASELINE-size inputs (20 kb, band 1000 / unbanded 5 kb): properties that hold at any size — identity
    alignment scores match*L with CIGAR LM, the global raw score is symmetric under swapping the sequences
    (test_cpp_wrappers.py:117-125), and the re-scored CIGAR equals the DP score."""
    rng = random.Random(5)
    a = ''.join(rng.choice('ACGT') for _ in range(20000))
    b = _mutate(a, 0.15, rng)
    r = ub.fully_global_alignment(a, a, SCHEME, True, 1000).split(',', 9)
    assert r[6] == str(3 * len(a)) and r[9] == '%dM' % len(a) and float(r[7]) == 100.0
    ab = ub.fully_global_alignment(a, b, SCHEME, True, 1000).split(',', 9)
    ba = ub.fully_global_alignment(b, a, SCHEME, True, 1000).split(',', 9)
    assert ab[6] == ba[6]
    assert (int(ab[2]), int(ab[3]), int(ab[4]), int(ab[5])) == (0, len(a), 0, len(b))
    a5, b5 = a[:5000], _mutate(a[:5000], 0.1, rng)
    u = ub.fully_global_alignment(a5, b5, SCHEME, False, 0).split(',', 9)
    w = ub.fully_global_alignment(a5, b5, SCHEME, True, 1000).split(',', 9)
    assert int(u[6]) >= int(w[6])
    stats = ub.last_stats()
    assert stats['launches'] >= 1 and stats['kernel_ms'] > 0


@pytest.mark.skipif(not os.path.isfile(REF_LIB), reason='oracle/_ref not built')
def test_against_reference_library_directly(ub):
    """When the unmodified reference library travelled with the snapshot, compare fresh random pairs."""
    from refdriver import AbiLib
    ref = AbiLib(REF_LIB)
    rng = random.Random(2024)
    s1s, s2s = [], []
    for _ in range(60):
        a = ''.join(rng.choice('ACGT') for _ in range(rng.randint(50, 1500)))
        s1s.append(a)
        s2s.append(_mutate(a, 0.2, rng))
    for banded, band in ((True, 100), (False, 0)):
        g = ub.fully_global_alignment_batch(s1s, s2s, SCHEME, banded, band)
        p = ub.path_alignment_batch(s1s, s2s, SCHEME, banded, band)
        for a, b, gg, pp in zip(s1s, s2s, g, p):
            assert mask_ms(gg) == mask_ms(ref.fully_global(a, b, SCHEME, banded, band))
            assert mask_ms(pp) == mask_ms(ref.path(a, b, SCHEME, banded, band))


def test_in_line_path_matches_two_pass_path(ub, monkeypatch):
    """The engine has two routes for a small chain grid: the two-pass route (score-only pass 1 + recorded pass 2)
    and the in-line route (full trace fill + traceback on the control warp, also the fallback).  Both must give the
    reference's strings."""
    d = load_golden('semiglobal_contained.json.gz')
    jobs = golden_chain_jobs(d)
    monkeypatch.setenv('UNICYCLER_B200_NO_FAST', '1')
    got = ub.chain_alignment_batch(jobs, tuple(d['scheme']), jobs[0]['band'])
    monkeypatch.delenv('UNICYCLER_B200_NO_FAST')
    bad = [(j['readName'], j['refName']) for j, g in zip(jobs, got) if mask_ms(g) != j['result']]
    assert not bad, bad[:5]


@pytest.mark.parametrize('seg_cycles', ['2.5e6', '6e6', '40e6'])
def test_result_does_not_depend_on_how_chains_are_cut_into_segments(ub, monkeypatch, seg_cycles):
    """Long seed chains are walked by several warps at once from guessed initialisation cells; the guess is only
    kept when it is proven right, so any segmentation (here: three different latency budgets per segment, from many
    short segments to none) must give the reference's strings.  Also guards the sizing of the task boards, which a
    grid enters once per segment that walks it."""
    d = load_golden('semiglobal_sample.json.gz')
    jobs = golden_chain_jobs(d)
    monkeypatch.setenv('UNICYCLER_B200_SEG_CYCLES', seg_cycles)
    got = ub.chain_alignment_batch(jobs, tuple(d['scheme']), jobs[0]['band'])
    monkeypatch.delenv('UNICYCLER_B200_SEG_CYCLES')
    bad = [(j['readName'], j['refName']) for j, g in zip(jobs, got) if mask_ms(g) != j['result']]
    assert not bad, bad[:5]


@pytest.mark.parametrize('switch', ['UNICYCLER_B200_NO_TILE_HELP', 'UNICYCLER_B200_NO_PERSIST', 'UNICYCLER_B200_NO_SPLIT'])
def test_scheduling_switches_do_not_change_results(ub, monkeypatch, switch):
    """Tile helpers (other warps recompute trace tiles ahead of a long traceback), persistent checkpoint blocks
    (big-grid tracebacks deferred to pass 2) and speculative segments only change WHEN work happens; with each of
    them switched off the sample_data chains must still give the reference's strings."""
    d = load_golden('semiglobal_sample.json.gz')
    jobs = golden_chain_jobs(d)
    monkeypatch.setenv(switch, '1')
    got = ub.chain_alignment_batch(jobs, tuple(d['scheme']), jobs[0]['band'])
    monkeypatch.delenv(switch)
    bad = [(j['readName'], j['refName']) for j, g in zip(jobs, got) if mask_ms(g) != j['result']]
    assert not bad, bad[:5]


@pytest.mark.skipif(not os.path.isfile(REF_LIB), reason='oracle/_ref not built')
def test_semi_global_with_ambiguous_and_lower_case_bases(ub):
    """k-mers that are not pure upper-case ACGT take the literal-string route of the k-mer index
    (src/kmers.cpp:51-65 keys by string); N matches N in the DP (Dna5).  Compared with the reference library."""
    from refdriver import AbiLib
    ref = AbiLib(REF_LIB)
    rng = random.Random(77)
    contig = ''.join(rng.choice('ACGT') for _ in range(6000))
    contig = contig[:1500] + 'NNNNNNNNNNNN' + contig[1512:3000] + contig[3000:3100].lower() + contig[3100:]
    reads = []
    for k in range(4):
        start = rng.randint(200, 2500)
        seq = _mutate(contig[start:start + 2500].upper(), 0.1, rng)
        if k % 2:
            seq = seq[:700] + 'NNN' + seq[703:]
        reads.append(('r%d' % k, seq, '0,%d,+,ctg,%d,%d' % (len(seq), start, start + 2500)))
    refs = [('ctg', contig)]
    h = ub.new_ref_seqs()
    ub.add_ref_seq(h, 'ctg', contig)
    out = ub.semi_global_alignment_batch([r[0] for r in reads], [r[1] for r in reads], [r[2] for r in reads], h, SCHEME, 0)
    hr = ref.new_refs(refs)
    for r, o in zip(reads, out):
        want = ref.semi_global(r[0], r[1], r[2], hr, SCHEME)
        assert mask_semi_global(o) == mask_semi_global(want), r[0]
    ref.delete_refs(hr)
    ub.delete_ref_seqs(h)
