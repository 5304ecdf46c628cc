"""Common k-mer join: the oracle restatement against the reference's own console counts,
the host join of the per-read entry point against the oracle (no GPU), and the device join of the batch path against
both (-m gpu)."""
import os
import random
import re
import sys

import pytest

from oracle_lib import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
from kmer_join_oracle import common_kmers, reverse_complement  # noqa: E402

K_OF_LEVEL = {0: 10, 1: 10, 2: 9, 3: 8}   # settings.h:17-42


def reference_console_ranges():
    """(read strand, window, k, count the reference printed) for every range in the verbosity-3 golden outputs."""
    g = load_golden('semiglobal_sensitivity.json.gz')
    cases = []
    for setname, entries in g['sets'].items():
        d = load_golden('semiglobal_%s.json.gz' % setname)
        reads = {r[0]: r for r in d['reads']}
        refs = dict(d['refs'])
        for e in entries:
            if e['verbosity'] < 3:
                continue
            console = e['expected'].split(';')[-1]
            # "Reference ranges:" lists name+strand in the order the "Range:" blocks follow (semi_global_align.cpp:57-63, 97-131)
            listed = re.findall(r'^    (\S+)([+-]): (\d+) - (\d+)$', console, re.M)
            blocks = re.findall(r'^Range: (\S+): (\d+) - (\d+)\n    common (\d+)-mers: (\d+)$', console, re.M)
            assert len(listed) == len(blocks) > 0
            seq = reads[e['read']][1]
            for (name, strand, s, t), (name2, s2, t2, k, count) in zip(listed, blocks):
                assert (name, s, t) == (name2, s2, t2) and int(k) == K_OF_LEVEL[e['sensitivity']]
                strand_seq = seq if strand == '+' else reverse_complement(seq)
                cases.append((strand_seq, refs[name], int(s), int(t) - int(s), int(k), int(count)))
    return cases


def awkward_cases():
    rng = random.Random(7)
    def rnd(n, alphabet='ACGT'):
        return ''.join(rng.choice(alphabet) for _ in range(n))
    ref = rnd(3000)
    read = ref[500:1500]
    cases = [
        (read, ref, 0, len(ref), 10),
        (read, ref, 400, 1300, 8),
        (read, ref, 400, 1300, 17),                                  # literal keys beyond 16 bases
        (read[:300] + 'N' * 12 + read[312:], ref[:700] + 'N' * 15 + ref[715:], 0, len(ref), 10),   # N runs on both sides
        (read[:200] + read[200:400].lower() + read[400:], ref[:800] + ref[800:900].lower() + ref[900:], 100, 2500, 9),
        ('A' * 400 + rnd(100), rnd(50) + 'A' * 300 + rnd(50), 0, 400, 10),   # one k-mer many times on both sides
        ('ACGT' * 100, 'ACGT' * 80, 3, 300, 10),                      # short-period repeats
        ('ACGTACGTA', ref, 0, 500, 10),                               # read shorter than k
        (read, ref, 10, 9, 10),                                       # window shorter than k
        (read, ref, 10, 10, 10),                                      # window of exactly one k-mer
        (rnd(2000, 'ACGTRYKMN'), rnd(2500, 'ACGTRYKMN'), 0, 2500, 8), # ambiguity codes as plain bytes
    ]
    return cases


def test_oracle_join_matches_the_reference_console_counts():
    cases = reference_console_ranges()
    assert len(cases) >= 20
    for strand_seq, ref, start, length, k, count in cases:
        assert len(common_kmers(strand_seq, ref[start:start + length], k)) == count


def test_host_join_matches_the_oracle(ub):
    cases = [c[:5] for c in reference_console_ranges()[:12]] + awkward_cases()
    for strand_seq, ref, start, length, k in cases:
        want = common_kmers(strand_seq, ref[start:start + length], k)
        assert ub.common_kmers(strand_seq, ref, start, length, k, on_device=False) == want


@pytest.mark.gpu
def test_device_join_matches_the_oracle_and_the_host_join(ub):
    cases = [c[:5] for c in reference_console_ranges()] + awkward_cases()
    for strand_seq, ref, start, length, k in cases:
        want = common_kmers(strand_seq, ref[start:start + length], k)
        got = ub.common_kmers(strand_seq, ref, start, length, k, on_device=True)
        assert got == want, (len(got), len(want), k, start, length)
        assert ub.common_kmers(strand_seq, ref, start, length, k, on_device=False) == want


@pytest.mark.gpu
def test_batch_path_joins_on_the_device_and_host_switch_gives_the_same_strings(ub, monkeypatch):
    """The batch ABI reports the launches / points of its device join; with the developer switch that keeps the
    host join the result strings are the same."""
    from oracle_lib import mask_semi_global
    d = load_golden('semiglobal_small.json.gz')
    h = ub.new_ref_seqs()
    for name, seq in d['refs']:
        ub.add_ref_seq(h, name, seq)
    reads = [r for r in d['reads'] if r[0] in d['expected']]
    args = ([r[0] for r in reads], [r[1] for r in reads], [r[2] for r in reads], h, tuple(d['scheme']), 0)
    out = ub.semi_global_alignment_batch(*args)
    js = ub.last_join_stats()
    assert js['launches'] >= 7 and js['points'] > 0 and js['kernel_ms'] > 0
    monkeypatch.setenv('UNICYCLER_B200_HOST_KMERS', '1')
    out_host = ub.semi_global_alignment_batch(*args)
    assert ub.last_join_stats()['launches'] == 0
    ub.delete_ref_seqs(h)
    assert [mask_semi_global(o) for o in out] == [mask_semi_global(o) for o in out_host]
    assert all(mask_semi_global(o) == d['expected'][r[0]] for r, o in zip(reads, out))


@pytest.mark.gpu
@pytest.mark.parametrize('setname,chunk', [('sample', 4), ('sample', 7), ('small', 1), ('tough', 3)])
def test_chunked_pipeline_gives_the_golden_strings(ub, monkeypatch, setname, chunk):
    """The pipeline of large batch calls (line tracing of chunk k+1 on the pool while chunk k is staged / launched /
    fetched / formatted, k-mer join two chunks ahead, engines alternating) forced onto a small set: same strings as
    the single-launch call and as the reference."""
    from oracle_lib import mask_semi_global
    d = load_golden('semiglobal_%s.json.gz' % setname)
    h = ub.new_ref_seqs()
    for name, seq in d['refs']:
        ub.add_ref_seq(h, name, seq)
    reads = [r for r in d['reads'] if r[0] in d['expected']]
    reads = reads + [('no_hits', reads[0][1], '')] if setname == 'small' else reads   # a read without minimap hits
    args = ([r[0] for r in reads], [r[1] for r in reads], [r[2] for r in reads], h, tuple(d['scheme']), 0)
    monkeypatch.setenv('UNICYCLER_B200_CHUNK_READS', str(chunk))
    out = ub.semi_global_alignment_batch(*args)
    assert ub.last_join_stats()['launches'] >= 14   # at least two chunks were joined on the device
    monkeypatch.delenv('UNICYCLER_B200_CHUNK_READS')
    whole = ub.semi_global_alignment_batch(*args)
    ub.delete_ref_seqs(h)
    assert [mask_semi_global(o) for o in out] == [mask_semi_global(o) for o in whole]
    for r, o in zip(reads, out):
        if r[0] in d['expected']:
            assert mask_semi_global(o) == d['expected'][r[0]], r[0]
        else:
            assert o == ''


@pytest.mark.gpu
def test_chunked_pipeline_over_all_devices_of_the_process():
    """UNICYCLER_B200_DEVICES=all: the chunks of one batch call go round every visible GPU (two engines each, the k-mer
    join on the first).  Read once per process, hence the subprocess; with one GPU this is the two-engine pipeline."""
    import subprocess
    code = '''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import unicycler_b200 as ub
from oracle_lib import load_golden, mask_semi_global
for setname, chunk in (("sample", "3"), ("tough", "2")):
    d = load_golden("semiglobal_%%s.json.gz" %% setname)
    h = ub.new_ref_seqs()
    for name, seq in d["refs"]:
        ub.add_ref_seq(h, name, seq)
    reads = [r for r in d["reads"] if r[0] in d["expected"]]
    os.environ["UNICYCLER_B200_CHUNK_READS"] = chunk
    out = ub.semi_global_alignment_batch([r[0] for r in reads], [r[1] for r in reads], [r[2] for r in reads], h, tuple(d["scheme"]), 0)
    ub.delete_ref_seqs(h)
    bad = [r[0] for r, o in zip(reads, out) if mask_semi_global(o) != d["expected"][r[0]]]
    assert not bad, (setname, bad)
    assert ub.last_join_stats()["launches"] >= 14
''' % (ROOT, os.path.join(ROOT, 'tests'))
    env = dict(os.environ, UNICYCLER_B200_DEVICES='all')
    r = subprocess.run([sys.executable, '-c', code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    assert r.returncode == 0, r.stdout.decode()[-2000:]


@pytest.mark.gpu
def test_concurrent_chunked_batch_calls(ub, monkeypatch):
    """Two threads in chunked batch calls at once (each such call holds several engines between stage and end; the
    library serialises them) while a third one makes per-read calls through the coalescer."""
    import threading
    from oracle_lib import mask_semi_global
    d = load_golden('semiglobal_sample.json.gz')
    h = ub.new_ref_seqs()
    for name, seq in d['refs']:
        ub.add_ref_seq(h, name, seq)
    reads = [r for r in d['reads'] if r[0] in d['expected']][:12]
    args = ([r[0] for r in reads], [r[1] for r in reads], [r[2] for r in reads], h, tuple(d['scheme']), 0)
    monkeypatch.setenv('UNICYCLER_B200_CHUNK_READS', '3')
    outs, errs = {}, []

    def batch(tag):
        try:
            outs[tag] = ub.semi_global_alignment_batch(*args)
        except Exception as e:   # noqa: BLE001
            errs.append(e)

    def single():
        try:
            m, mm, go, ge = tuple(d['scheme'])
            outs['single'] = [ub.semi_global_alignment(r[0], r[1], 0, r[2], h, m, mm, go, ge, 0.0, False, 0) for r in reads[:4]]
        except Exception as e:   # noqa: BLE001
            errs.append(e)

    threads = [threading.Thread(target=batch, args=('a',)), threading.Thread(target=batch, args=('b',)),
               threading.Thread(target=single)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
        assert not t.is_alive(), 'batch calls wait for each other'
    ub.delete_ref_seqs(h)
    assert not errs, errs
    want = [d['expected'][r[0]] for r in reads]
    assert [mask_semi_global(o) for o in outs['a']] == want
    assert [mask_semi_global(o) for o in outs['b']] == want
    assert [mask_semi_global(o) for o in outs['single']] == want[:4]
