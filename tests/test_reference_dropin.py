"""Drop-in tests: libunicycler_b200.so installed as unicycler/cpp_functions.so inside the reference's OWN, unmodified
Python package (staged by oracle/Makefile.ref into oracle/_ref/pydist — a build output that travels to the GPU box,
like oracle/_ref/libunicycler_ref.so), with UNICYCLER_B200_FORWARD_LIB pointing at the reference library for the
symbols outside the hot path (minimap, miniasm, consensus).

  * the reference's unit tests for this path, test/test_cpp_wrappers.py and test/test_semi_global_alignment.py
    (49 tests; they drive unicycler_align.semi_global_align_long_reads with its thread pool, i.e. the per-read ABI
    behind the request coalescer);
  * bridge path scoring: path_finding.get_best_paths_for_seq (which the reference never tests) on
    test/test_assembly_graph.gfa, per-call and with the batch prefetch of INTEGRATION.md, against the paths and scores
    the reference library produced (tests/golden/bridge_tuples.json.gz);
  * non-GPU: every forwarder resolves and returns the reference's value.
"""
import ctypes
import json
import os
import shutil
import subprocess
import sys

import pytest

from oracle_lib import REF_LIB, ROOT, load_golden, mask_ms

PYDIST = os.path.join(ROOT, 'oracle', '_ref', 'pydist')
needs_ref = pytest.mark.skipif(not (os.path.isfile(REF_LIB) and os.path.isdir(PYDIST)),
                               reason='oracle/_ref (reference library + staged Python package) not built')


def _stage(tmp_path, ub):
    work = str(tmp_path / 'ref_pkg')
    shutil.copytree(PYDIST, work)
    shutil.copy(ub.LIB_PATH, os.path.join(work, 'unicycler', 'cpp_functions.so'))
    env = dict(os.environ, UNICYCLER_B200_FORWARD_LIB=REF_LIB, PYTHONWARNINGS='ignore')
    return work, env


@needs_ref
@pytest.mark.gpu
def test_reference_unit_tests_pass_on_the_replacement_library(ub, tmp_path):
    work, env = _stage(tmp_path, ub)
    r = subprocess.run([sys.executable, '-m', 'unittest', 'test.test_cpp_wrappers', 'test.test_semi_global_alignment'],
                       cwd=work, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=1500)
    text = r.stdout.decode()
    assert r.returncode == 0, text[-3000:]
    assert 'Ran 49 tests' in text and 'OK' in text, text[-1500:]


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize('mode', ['direct', 'prefetch'])
def test_bridge_path_scoring_through_reference_python(ub, tmp_path, mode):
    work, env = _stage(tmp_path, ub)
    out = os.path.join(work, 'bridge.json')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'dropin_bridge_harness.py'), work, out, mode, '24'],
                       cwd=work, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=1500)
    assert r.returncode == 0, r.stdout.decode()[-3000:]
    got = json.load(open(out))
    want = load_golden('bridge_tuples.json.gz')
    assert len(got['bridges']) == len(want['bridges'])
    for g, w in zip(got['bridges'], want['bridges']):
        assert (g['start'], g['end'], g['true_path']) == (w['start'], w['end'], w['true_path'])
        assert g['result'] == w['result'], (g['start'], g['end'])          # same paths, raw and scaled scores, order
        assert g['progressive'] == w['progressive'] and g['calls'] == w['calls']
    strings = want['strings']
    assert len(got['recorded']) == len(want['recorded'])
    for a, b in zip(got['recorded'], want['recorded']):
        assert (a['fn'], a['s1'], a['s2'], a['band']) == (b['fn'], strings[b['s1']], strings[b['s2']], b['band'])
        assert a['result'] == b['result']


@needs_ref
@pytest.mark.gpu
def test_reference_alignment_driver_with_thread_pool(ub, tmp_path):
    """unicycler_align.semi_global_align_long_reads, unmodified, with 8 Python threads on the sample_data reads: the
    per-read ABI behind the request coalescer.  Same alignments kept, same coordinates, scores and CIGARs as with the
    reference library (tests/golden/dropin_align_sample.json.gz)."""
    work, env = _stage(tmp_path, ub)
    out = os.path.join(work, 'align.json')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'dropin_align_harness.py'), work, out, '8'],
                       cwd=work, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=1500)
    assert r.returncode == 0, r.stdout.decode()[-3000:]
    got = json.load(open(out))
    want = load_golden('dropin_align_sample.json.gz')
    assert got['reads'] == want['reads']
    assert sum(len(v) for v in got['reads'].values()) >= 30


@pytest.mark.gpu
def test_bridge_tuples_golden_batch(ub):
    """The alignments of the bridge run as device batches (config-3 benchmark input)."""
    d = load_golden('bridge_tuples.json.gz')
    strings, sc = d['strings'], tuple(d['scheme'])
    groups = {}
    for r in d['recorded']:
        groups.setdefault((r['fn'], r['banded'], r['band']), []).append(r)
    n = 0
    for (fn, banded, band), rs in groups.items():
        f = ub.fully_global_alignment_batch if fn == 'global' else ub.path_alignment_batch
        out = f([strings[r['s1']] for r in rs], [strings[r['s2']] for r in rs], sc, banded, band)
        for r, o in zip(rs, out):
            assert mask_ms(o) == r['result'], (fn, band, len(strings[r['s1']]), len(strings[r['s2']]))
            n += 1
    assert n == len(d['recorded']) and n > 200


@needs_ref
def test_forwarders_reach_the_reference_library(ub, tmp_path):
    """Symbols outside the accelerated path are forwarded with the reference's exact C types (no GPU needed)."""
    code = r'''
import ctypes, os, sys
lib = ctypes.CDLL(sys.argv[1]); ref = ctypes.CDLL(sys.argv[2])
for L in (lib, ref):
    L.freeCString.argtypes = [ctypes.c_void_p]
    L.minimapAlignReads.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.minimapAlignReads.restype = ctypes.c_void_p
    L.simulateDepths.argtypes = [ctypes.POINTER(ctypes.c_int)] + [ctypes.c_int] * 4
    L.simulateDepths.restype = ctypes.c_void_p
    L.multipleSequenceAlignment.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p),
                                            ctypes.c_ulong, ctypes.c_uint] + [ctypes.c_int] * 4
    L.multipleSequenceAlignment.restype = ctypes.c_void_p
def s(L, p):
    v = ctypes.cast(p, ctypes.c_char_p).value.decode(); L.freeCString(p); return v
fa, fq = sys.argv[3].encode(), sys.argv[4].encode()
a = s(lib, lib.minimapAlignReads(fa, fq, 1, 0, 0)); b = s(ref, ref.minimapAlignReads(fa, fq, 1, 0, 0))
assert a == b and len(a) > 100, (len(a), len(b))
seqs = [b'ACGTTGCAAGCTTGCATGCCTGCAGGTCGACTCTAGAGGATCC', b'ACGTTGCAAGCTTGCATCCTGCAGGTCGACTCTAGAGGATCC',
        b'ACGTTGCAAGCTTGCATGCCTGCAGGTCGACTTTAGAGGATCC']
quals = [b'I' * len(x) for x in seqs]
arr = (ctypes.c_char_p * 3)(*seqs); q = (ctypes.c_char_p * 3)(*quals)
a = s(lib, lib.multipleSequenceAlignment(arr, q, 3, 1000, 3, -6, -5, -2))
b = s(ref, ref.multipleSequenceAlignment(arr, q, 3, 1000, 3, -6, -5, -2))
assert a == b and a, (a, b)
print('forwarders ok')
'''
    fa = os.path.join(PYDIST, 'test', 'test_semi_global_alignment.fasta')
    fq = os.path.join(PYDIST, 'test', 'test_semi_global_alignment.fastq')
    env = dict(os.environ, UNICYCLER_B200_FORWARD_LIB=REF_LIB)
    r = subprocess.run([sys.executable, '-c', code, ub.LIB_PATH, REF_LIB, fa, fq], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, timeout=300)
    assert r.returncode == 0 and 'forwarders ok' in r.stdout.decode(), r.stdout.decode()[-2000:]


def test_start_end_alignment_return_int(ub):
    """startAlignment / endAlignment return int like include/start_end_align.h:22-24 (cpp_wrappers.py:336,352)."""
    header = open(os.path.join(ROOT, 'include', 'unicycler_b200.h')).read()
    assert 'int startAlignment(' in header and 'int endAlignment(' in header
    lib = ctypes.CDLL(ub.LIB_PATH)
    assert lib.startAlignment and lib.endAlignment


@needs_ref
def test_alignment_tallies_match_the_reference_python(ub, tmp_path):
    """ub200_alignmentTallies against Alignment.tally_up_score_and_errors of the reference's own
    alignment.py (run in a subprocess inside the staged package) on every golden alignment of two fixtures.  No GPU."""
    work = str(tmp_path / 'ref_pkg')
    shutil.copytree(PYDIST, work)
    shutil.copy(REF_LIB, os.path.join(work, 'unicycler', 'cpp_functions.so'))
    code = r'''
import gzip, json, sys
sys.path.insert(0, sys.argv[1])
import unicycler.alignment, unicycler.read_ref
scheme = unicycler.alignment.AlignmentScoringScheme('3,-6,-5,-2')
out = []
for name in ('small', 'sample'):
    d = json.load(gzip.open(sys.argv[2] + '/semiglobal_%s.json.gz' % name, 'rt'))
    refs = {n: unicycler.read_ref.Reference(n, s) for n, s in d['refs']}
    for rn, seq, _ in d['reads']:
        read = unicycler.read_ref.Read(rn, seq, 'I' * len(seq))
        for k, s in enumerate(d['expected'].get(rn, '').split(';')[:-1]):
            a = unicycler.alignment.Alignment(seqan_output=s, read=read, reference_dict=refs, scoring_scheme=scheme)
            out.append([name, rn, k, a.match_count, a.mismatch_count, a.insertion_count, a.deletion_count,
                        a.raw_score, a.alignment_length, repr(a.percent_identity), repr(a.scaled_score), a.edit_distance])
json.dump(out, open(sys.argv[3], 'w'))
'''
    res = os.path.join(work, 'tallies.json')
    r = subprocess.run([sys.executable, '-W', 'ignore', '-c', code, work, os.path.join(ROOT, 'tests', 'golden'), res],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert r.returncode == 0, r.stdout.decode()[-2000:]
    want = json.load(open(res))
    from oracle_lib import revcomp
    data = {}
    for name in ('small', 'sample'):
        d = load_golden('semiglobal_%s.json.gz' % name)
        data[name] = (dict(d['refs']), {r[0]: r[1] for r in d['reads']}, d)
    n = 0
    for name, rn, k, mc, mmc, ic, dc, raw, alen, ident, scaled, edit in want:
        refs, reads, d = data[name]
        s = d['expected'][rn].split(';')[k].split(',', 9)
        read_seq = reads[rn] if s[1] == '+' else revcomp(reads[rn])
        t = ub.alignment_tallies(read_seq, refs[s[0]], int(s[2]), int(s[4]), s[9], (3, -6, -5, -2))
        assert (t['match_count'], t['mismatch_count'], t['insertion_count'], t['deletion_count'], t['raw_score'],
                t['alignment_length'], t['edit_distance']) == (mc, mmc, ic, dc, raw, alen, edit), (name, rn)
        assert repr(t['percent_identity']) == ident and repr(t['scaled_score']) == scaled, (name, rn)
        n += 1
    assert n >= 40
