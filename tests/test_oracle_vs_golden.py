"""Pins the oracle (oracle/dp_oracle.cpp) against golden vectors produced by the UNMODIFIED reference library
(tests/golden/make_golden.py): the reference's own known-answer inputs plus differential cases.  CPU only."""
import pytest

from oracle_lib import golden_chain_jobs, load_golden, mask_ms


def test_global_and_path_match_reference(oracle):
    d = load_golden('global_path.json.gz')
    assert len(d['cases']) > 600
    for c in d['cases']:
        sc = tuple(c['scheme'])
        assert mask_ms(oracle.fully_global(c['s1'], c['s2'], sc, c['banded'], c['band'])) == c['global']
        assert mask_ms(oracle.path(c['s1'], c['s2'], sc, c['banded'], c['band'])) == c['path']


def test_reference_known_answers_global():
    """test/test_cpp_wrappers.py:22-125 pins raw scores for hand-made pairs; the golden file carries the
    reference's full strings for the same FASTA, and the exact known scores are re-asserted here."""
    d = load_golden('global_path.json.gz')
    kats = [c for c in d['cases'] if c['kind'] == 'kat']
    assert kats
    for c in kats:
        assert c['global'].split(',')[0] == 's2'
        if c['s1'] == c['s2'] and not c['banded']:
            assert int(c['global'].split(',')[6]) == 3 * len(c['s1'])


@pytest.mark.parametrize('setname', ['small', 'contained'])
def test_chain_matches_reference(oracle, setname):
    d = load_golden('semiglobal_%s.json.gz' % setname)
    jobs = golden_chain_jobs(d)
    assert jobs
    for j in jobs:
        got = mask_ms(oracle.chain(j['readSeq'], j['refSeq'], j['seeds'], tuple(d['scheme']), j['band'], j['readName'],
                                   j['refName'], j['refOffset']))
        assert got == j['result'], j['readName']


def test_chain_tough_subset(oracle):
    d = load_golden('semiglobal_tough.json.gz')
    jobs = [j for j in golden_chain_jobs(d) if len(j['readSeq']) < 12000][:20]
    assert len(jobs) >= 10
    for j in jobs:
        got = mask_ms(oracle.chain(j['readSeq'], j['refSeq'], j['seeds'], tuple(d['scheme']), j['band'], j['readName'],
                                   j['refName'], j['refOffset']))
        assert got == j['result'], j['readName']


def test_perfect_match_known_answers():
    """test/test_semi_global_alignment.py:22-227: exact score / coordinates / CIGAR of the 9 synthetic reads."""
    d = load_golden('semiglobal_small.json.gz')
    want = {'0': ('0', 0, 100, 60, 160, 300, '100M'), '1': ('1', 0, 200, 100, 300, 600, '200M'),
            '2': ('2', 0, 150, 0, 150, 450, '150M'), '3': ('3', 62, 162, 0, 100, 300, '62S100M138S'),
            '4': ('4', 0, 120, 0, 120, 360, '120M180S'), '5': ('5', 120, 300, 0, 180, 540, '120S180M'),
            '6': ('6', 190, 300, 0, 110, 330, '190S110M'), '7': ('7', 0, 130, 170, 300, 390, '130M170S'),
            '8': ('8', 0, 300, 0, 300, 900, '300M')}
    for name, exp in want.items():
        parts = d['expected'][name].split(';')[0].split(',', 9)
        got = (parts[0], int(parts[2]), int(parts[3]), int(parts[4]), int(parts[5]), int(parts[6]), parts[9])
        assert got == exp
        assert float(parts[7]) == 100.0
