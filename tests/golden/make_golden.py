#!/usr/bin/env python3
"""Generates the golden vectors under tests/golden/ from the UNMODIFIED reference library
(oracle/_ref/libunicycler_ref.so, built from /root/reference by oracle/Makefile.ref).

Run in the build container only (it needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden.py

Outputs (all gzip-compressed JSON):
  global_path.json.gz        fullyGlobalAlignment / pathAlignment cases: the reference's own known-answer
                             inputs (test/test_cpp_wrappers.fasta, test_cpp_wrappers.py:22-125) plus seeded
                             random pairs; expected = reference output with the milliseconds field masked.
  freeend.json.gz            semiGlobalAlignmentExhaustive / startAlignment / endAlignment / overlapAlignment cases
                              with the reference's return values.
  pairs_large.json.gz        BASELINE configs[2]-sized pairs (5 / 10 / 20 kb, bands 1000 / 500 / 50, both entry points).
  calibration_pairs.json.gz  getRandomSequenceAlignmentScores pinned per pair: the pairs the product generates for a
                             fixed seed (ub200_calibrationPairs; same std::mt19937 stream as random_alignments.cpp:33-40)
                             through the reference's fullyGlobalAlignment, plus the "mean,sd" string computed from the
                             reference's scaled scores with getMeanAndStDev's arithmetic (random_alignments.cpp:187-202).
  semiglobal_sensitivity.json.gz  semiGlobalAlignment at sensitivity levels 1-3 (settings.h:17-42) and with
                             verbosity 3 (console text) on reads of the small / sample / contained sets.
  semiglobal_synth5.json.gz  BASELINE configs[4]-style inputs (random 1 Mbp reference, 20 kb reads with 15 % errors, hit
                             strings from the ground truth); sequences are regenerated from the seed, only the
                             expected strings are stored.
  bridge_tuples.json.gz      bridge path scoring (BASELINE configs[2], SURVEY.md 8(c) substitute): the reference's own
                             path_finding.get_best_paths_for_seq on test/test_assembly_graph.gfa with synthetic read
                             consensus sequences (tests/dropin_bridge_harness.py): chosen paths and every
                             (s1, s2, band, entry point) that crossed the seam with the reference's result.
  dropin_align_sample.json.gz  what unicycler_align.semi_global_align_long_reads (the reference's Python driver, thread
                             pool of 8) keeps for every sample_data read with the reference library underneath.
  semiglobal_<set>.json.gz   for the reference's semi-global fixtures (test/test_semi_global_alignment*.{fasta,fastq}
                             and sample_data): references, reads, the minimap hit strings the reference's own
                             minimap produces, the expected semiGlobalAlignment output per read, and — from a
                             scratch build of the reference with a dump hook in alignReadToReferenceRange (patch
                             below, applied to a copy under /tmp, never to /root/reference) — the seed chain of
                             every bandedChainAlignment call with its result.  The hook build's results are
                             checked to be identical to the unmodified library's before anything is written.
"""
import ctypes
import glob
import gzip
import json
import os
import random
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from refdriver import AbiLib, load_fasta, load_fastq, mask_ms, mask_semi_global  # noqa: E402

REF = os.environ.get('UNICYCLER_REFERENCE', '/root/reference')
REF_LIB = os.path.join(ROOT, 'oracle', '_ref', 'libunicycler_ref.so')
INSTR_DIR = '/tmp/unicycler_instr'
SCHEME = (3, -6, -5, -2)

DUMP_PATCH_MARKER = "        // Finally we can actually do the Seqan alignment!"
DUMP_CODE = r'''
        FILE * dumpF = 0;
        if (getenv("UNICYCLER_DUMP")) dumpF = fopen(getenv("UNICYCLER_DUMP"), "a");
        if (dumpF) {
            fprintf(dumpF, "JOB\t%s%c\t%s\t%d\t%d\t%d\t%d\t%d\t%d\t%d\n", readName.c_str(), readStrand, refName.c_str(), refStart, bandSize, matchScore, mismatchScore, gapOpenScore, gapExtensionScore, (int)length(seedChain));
            fprintf(dumpF, "LEN\t%d\t%d\n", (int)readSeq->length(), (int)trimmedRefSeq.length());
            for (unsigned si = 0; si < length(seedChain); ++si)
                fprintf(dumpF, "SEED\t%ld\t%ld\t%ld\t%ld\t%ld\t%ld\n", (long)beginPositionH(seedChain[si]), (long)beginPositionV(seedChain[si]), (long)endPositionH(seedChain[si]), (long)endPositionV(seedChain[si]), (long)lowerDiagonal(seedChain[si]), (long)upperDiagonal(seedChain[si]));
        }
'''


def build_instrumented():
    """Scratch copy of semi_global_align.cpp with a seed-chain dump hook; linked against the objects of the
    unmodified build.  Debug aid for fixture generation only."""
    os.makedirs(INSTR_DIR, exist_ok=True)
    src = open(os.path.join(REF, 'unicycler', 'src', 'semi_global_align.cpp')).read()
    assert DUMP_PATCH_MARKER in src
    src = src.replace(DUMP_PATCH_MARKER, DUMP_CODE + DUMP_PATCH_MARKER)
    old = "            alignments.push_back(sgAlignment);\n        }\n        catch (...) {}"
    assert old in src
    src = src.replace(old,
                      "            alignments.push_back(sgAlignment);\n"
                      "            if (dumpF) { fprintf(dumpF, \"RESULT\\t%s\\n\", sgAlignment->getFullString().c_str()); }\n"
                      "        }\n        catch (...) { if (dumpF) fprintf(dumpF, \"RESULT\\t\\n\"); }\n"
                      "        if (dumpF) fclose(dumpF);")
    src = src.replace('#include "settings.h"', '#include "settings.h"\n#include <cstdio>\n#include <cstdlib>')
    cpp = os.path.join(INSTR_DIR, 'semi_global_align.cpp')
    open(cpp, 'w').write(src)
    obj = os.path.join(INSTR_DIR, 'semi_global_align.o')
    lib = os.path.join(INSTR_DIR, 'libref_instr.so')
    subprocess.check_call(['g++', '-std=c++14', '-O3', '-DNDEBUG', '-fPIC', '-w',
                           '-I' + os.path.join(REF, 'unicycler', 'include'), '-c', '-o', obj, cpp])
    objs = [o for o in glob.glob(os.path.join(ROOT, 'oracle', '_ref', 'obj', '**', '*.o'), recursive=True)
            if not o.endswith('/semi_global_align.o')]
    subprocess.check_call(['g++', '-shared', '-o', lib, obj] + objs + ['-lz', '-lpthread'])
    return lib


def parse_dump(path):
    jobs, job = [], None
    for line in open(path):
        p = line.rstrip('\n').split('\t')
        if p[0] == 'JOB':
            job = dict(read=p[1], ref=p[2], refStart=int(p[3]), band=int(p[4]), seeds=[])
        elif p[0] == 'LEN':
            job['readLen'], job['refLen'] = int(p[1]), int(p[2])
        elif p[0] == 'SEED':
            job['seeds'].append([int(x) for x in p[1:7]])
        elif p[0] == 'RESULT':
            job['result'] = mask_ms(p[1]) if len(p) > 1 else ''
            jobs.append(job)
    return jobs


def write_json_gz(name, obj):
    path = os.path.join(HERE, name)
    with gzip.open(path, 'wt', compresslevel=9) as f:
        json.dump(obj, f, separators=(',', ':'))
    print('wrote', path, os.path.getsize(path), 'bytes')


def make_semiglobal(setname, ref_fa, reads_fq, ref_lib, instr_lib_path):
    refs = load_fasta(ref_fa)
    reads = load_fastq(reads_fq)
    fq = reads_fq
    if reads_fq.endswith('.gz'):
        fq = '/tmp/_golden_reads.fastq'
        with open(fq, 'w') as f:
            for n, s in reads:
                f.write('@%s\n%s\n+\n%s\n' % (n, s, 'I' * len(s)))
    hits = ref_lib.minimap_hits(ref_fa, fq)
    h = ref_lib.new_refs(refs)
    expected = {}
    for name, seq in reads:
        if name in hits:
            expected[name] = mask_semi_global(ref_lib.semi_global(name, seq, hits[name], h, SCHEME))
    ref_lib.delete_refs(h)
    # seed chains from the hook build (same inputs), cross-checked
    dump = os.path.join(INSTR_DIR, 'dump_%s.txt' % setname)
    if os.path.exists(dump):
        os.remove(dump)
    os.environ['UNICYCLER_DUMP'] = dump
    instr = AbiLib(instr_lib_path)
    h2 = instr.new_refs(refs)
    for name, seq in reads:
        if name in hits:
            out = mask_semi_global(instr.semi_global(name, seq, hits[name], h2, SCHEME))
            assert out == expected[name], 'hook build differs from the unmodified reference for read ' + name
    instr.delete_refs(h2)
    del os.environ['UNICYCLER_DUMP']
    jobs = parse_dump(dump)
    obj = dict(set=setname, scheme=list(SCHEME), sensitivity=0, refs=refs,
               reads=[[n, s, hits.get(n, '')] for n, s in reads], expected=expected, jobs=jobs)
    write_json_gz('semiglobal_%s.json.gz' % setname, obj)


def mutate(s, rate, rng):
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
    return ''.join(out) or 'A'


def make_global_path(ref_lib):
    cases = []
    # the reference's own known-answer inputs (test/test_cpp_wrappers.py:22-125)
    seqs = dict(load_fasta(os.path.join(REF, 'test', 'test_cpp_wrappers.fasta')))
    names = sorted(seqs)
    kat_pairs = [(a, b) for a in names for b in names if a < b and abs(len(seqs[a]) - len(seqs[b])) < 200][:40]
    for a, b in kat_pairs:
        for banded, band in ((False, 0), (True, 1000), (True, 10)):
            cases.append(dict(kind='kat', s1=seqs[a], s2=seqs[b], scheme=list(SCHEME), banded=banded, band=band))
    rng = random.Random(20240607)
    schemes = [(3, -6, -5, -2), (1, -1, -1, -1), (5, -4, -8, -6), (1, -3, -5, -2), (2, -2, -2, -2), (3, -6, -2, -5)]
    for it in range(600):
        L = rng.choice([2, 3, 5, 8, 20, 50, 100, 300, 700, 1500])
        s1 = ''.join(rng.choice('ACGTN' if rng.random() < 0.03 else 'ACGT') for _ in range(rng.randint(2, L)))
        mode = rng.random()
        if mode < 0.6:
            s2 = mutate(s1, rng.choice([0, 0.05, 0.15, 0.4]), rng)
        elif mode < 0.8:
            s2 = mutate(s1, 0.1, rng) + ''.join(rng.choice('ACGT') for _ in range(rng.randint(0, L)))
        else:
            s2 = ''.join(rng.choice('ACGT') for _ in range(rng.randint(2, L)))
        if len(s2) < 2:
            s2 += 'AC'
        if rng.random() < 0.3:
            s1, s2 = s2, s1
        banded = rng.random() < 0.6
        cases.append(dict(kind='rand', s1=s1, s2=s2, scheme=list(rng.choice(schemes)), banded=banded,
                          band=rng.choice([2, 3, 5, 10, 50, 500, 1000])))
    for c in cases:
        sc = tuple(c['scheme'])
        c['global'] = mask_ms(ref_lib.fully_global(c['s1'], c['s2'], sc, c['banded'], c['band']))
        c['path'] = mask_ms(ref_lib.path(c['s1'], c['s2'], sc, c['banded'], c['band']))
    write_json_gz('global_path.json.gz', dict(cases=cases))


def rand_seq(rng, n):
    """Reproducible across Python versions: only Random.random() is guaranteed stable."""
    return ''.join('ACGT'[int(rng.random() * 4)] for _ in range(n))


def noisy(s, rng, sub=0.05, dele=0.05, ins=0.05):
    out = []
    for c in s:
        r = rng.random()
        if r < sub:
            out.append('ACGT'[int(rng.random() * 4)])
        elif r < sub + dele:
            continue
        elif r < sub + dele + ins:
            out.append(c)
            out.append('ACGT'[int(rng.random() * 4)])
        else:
            out.append(c)
    return ''.join(out)


def make_freeend(ref_lib):
    rng = random.Random(8101)
    schemes = [(3, -6, -5, -2), (1, -1, -1, -1), (5, -4, -8, -6), (3, -6, -2, -5)]
    cases = []
    for it in range(160):
        sc = list(schemes[it % len(schemes)]) if it >= 80 else list(SCHEME)
        kind = ('exhaustive', 'start', 'end', 'overlap')[it % 4]
        if kind == 'exhaustive':
            # string_graph.py:323-328: a short sequence end against a longer window
            a = rand_seq(rng, rng.choice([12, 40, 100, 300, 700]))
            b = rand_seq(rng, rng.randint(0, 400)) + noisy(a, rng, 0.03, 0.03, 0.03) + rand_seq(rng, rng.randint(0, 400))
            if rng.random() < 0.3:
                a, b = b, a
            c = dict(kind=kind, s1=a, s2=b, scheme=sc)
            c['result'] = mask_ms(ref_lib.exhaustive(a, b, tuple(sc)))
        elif kind in ('start', 'end'):
            # miniasm_assembly.py:550-554: s1 expected at the start / end of s2 (s2 is trimmed to 1.5 x |s1|)
            a = rand_seq(rng, rng.choice([20, 60, 150, 400, 900]))
            tail = rand_seq(rng, rng.choice([0, 10, 200, 2000]))
            b = noisy(a, rng, 0.04, 0.04, 0.04)
            b = (b + tail) if kind == 'start' else (tail + b)
            if rng.random() < 0.15:
                b = rand_seq(rng, len(b))  # unrelated
            c = dict(kind=kind, s1=a, s2=b, scheme=sc)
            c['result'] = (ref_lib.start if kind == 'start' else ref_lib.end)(a, b, tuple(sc))
        else:
            ov = rng.choice([0, 5, 25, 77, 127, 400])
            core = rand_seq(rng, ov)
            a = rand_seq(rng, rng.randint(50, 900)) + core
            b = noisy(core, rng, 0.02, 0.02, 0.02) + rand_seq(rng, rng.randint(50, 900))
            guess = max(0, ov + rng.randint(-20, 20))
            c = dict(kind=kind, s1=a, s2=b, scheme=sc, guess=guess)
            c['result'] = ref_lib.overlap(a, b, tuple(sc), guess)
        cases.append(c)
    write_json_gz('freeend.json.gz', dict(cases=cases))


def make_pairs_large(ref_lib):
    """Bridge path scoring sizes (path_finding.py:71,323,334 band 1000 / 500; bridge_long_read_simple.py:485 band 50)."""
    from multiprocessing.dummy import Pool as ThreadPool
    rng = random.Random(777)
    cases = []
    for L, div in ((20000, 0.12), (20000, 0.04), (10000, 0.15), (10000, 0.10), (5000, 0.12), (5000, 0.25)):
        a = rand_seq(rng, L)
        b = noisy(a, rng, div / 3, div / 3, div / 3)
        if rng.random() < 0.5:
            b = b[:int(len(b) * rng.uniform(0.5, 0.95))]   # a partial path (pathAlignment's use)
        for band in (1000, 500, 50):
            cases.append(dict(s1=a, s2=b, scheme=list(SCHEME), banded=True, band=band))
    # one unbanded 5 kb pair (what the calibration does at L = 5000)
    a = rand_seq(rng, 5000)
    cases.append(dict(s1=a, s2=rand_seq(rng, 5000), scheme=list(SCHEME), banded=False, band=0))

    def one(c):
        sc = tuple(c['scheme'])
        c['global'] = mask_ms(ref_lib.fully_global(c['s1'], c['s2'], sc, c['banded'], c['band']))
        c['path'] = mask_ms(ref_lib.path(c['s1'], c['s2'], sc, c['banded'], c['band']))

    pool = ThreadPool(8)
    pool.map(one, cases)
    pool.close()
    write_json_gz('pairs_large.json.gz', dict(cases=cases))


def make_calibration(ref_lib):
    from multiprocessing.dummy import Pool as ThreadPool
    sys.path.insert(0, ROOT)
    import unicycler_b200 as ub   # sequence generator only (no GPU needed)
    sets = []
    for L, n, seed in ((100, 400, 42), (1000, 24, 42), (5000, 8, 7)):
        s1, s2 = ub.calibration_pairs(L, n, seed)
        pool = ThreadPool(8)
        res = pool.map(lambda ab: mask_ms(ref_lib.fully_global(ab[0], ab[1], SCHEME, False, 0)), list(zip(s1, s2)))
        pool.close()
        scores = [float(r.split(',')[7]) for r in res if r]
        # getMeanAndStDev (random_alignments.cpp:187-202) on the reference's scaled scores.  The "%f" field has the
        # full double only to 6 decimals, so the expected string is re-derived from exact integers below.
        exact = []
        for r in res:
            f = r.split(',', 9)
            raw = int(f[6])
            cols = sum(int(x) for x in __import__('re').findall(r'(\d+)[MID]', f[9]))
            perfect, worst = SCHEME[0] * cols, SCHEME[1] * cols
            exact.append(100.0 * float(raw - worst) / float(perfect - worst))
        mean = 0.0
        for v in exact:
            mean += v
        mean /= len(exact)
        dev = 0.0
        for v in exact:
            d = v - mean
            dev += d * d
        sd = __import__("math").sqrt(dev / len(exact))
        assert all(abs(a - b) < 1e-6 for a, b in zip(scores, exact))
        sets.append(dict(length=L, n=n, seed=seed, scheme=list(SCHEME), results=res, mean_sd='%f,%f' % (mean, sd),
                         first_pair=[s1[0], s2[0]]))
    write_json_gz('calibration_pairs.json.gz', dict(sets=sets))


def synth5_inputs(ref_len=1000000, n_reads=16, read_len=20000, seed=5):
    """BASELINE configs[4]-style inputs, regenerated identically by the tests (Random.random() only)."""
    rng = random.Random(seed)
    ref = rand_seq(rng, ref_len)
    comp = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A'}
    reads = []
    for k in range(n_reads):
        L = int(read_len * (0.9 + 0.2 * rng.random()))
        start = int(rng.random() * (ref_len - L))
        frag = ref[start:start + L]
        strand = '+' if rng.random() < 0.5 else '-'
        if strand == '-':
            frag = ''.join(comp[c] for c in reversed(frag))
        seq = noisy(frag, rng)
        reads.append(['read%d' % k, seq, '0,%d,%s,ref,%d,%d' % (len(seq), strand, start, start + L)])
    return ref, reads


def make_synth5(ref_lib):
    import hashlib
    from multiprocessing.dummy import Pool as ThreadPool
    ref, reads = synth5_inputs()
    h = ref_lib.new_refs([('ref', ref)])
    pool = ThreadPool(8)
    outs = pool.map(lambda r: mask_semi_global(ref_lib.semi_global(r[0], r[1], r[2], h, SCHEME)), reads)
    pool.close()
    ref_lib.delete_refs(h)
    write_json_gz('semiglobal_synth5.json.gz',
                  dict(ref_len=len(ref), n_reads=len(reads), read_len=20000, seed=5, scheme=list(SCHEME), sensitivity=0,
                       ref_sha1=hashlib.sha1(ref.encode()).hexdigest(),
                       reads_sha1=hashlib.sha1(''.join(r[1] for r in reads).encode()).hexdigest(),
                       expected={r[0]: o for r, o in zip(reads, outs)}))


def make_sensitivity(ref_lib):
    """Sensitivity levels 1-3 and the verbosity-3 console text, on reads of the committed fixtures."""
    from multiprocessing.dummy import Pool as ThreadPool
    import gzip as _gz
    out = dict(scheme=list(SCHEME), sets={})
    picks = dict(small=None, sample=8, contained=2)
    for setname, limit in picks.items():
        with _gz.open(os.path.join(HERE, 'semiglobal_%s.json.gz' % setname), 'rt') as f:
            d = json.load(f)
        reads = [r for r in d['reads'] if r[0] in d['expected']]
        if limit is not None:
            reads = sorted(reads, key=lambda r: len(r[1]))[:limit]
        h = ref_lib.new_refs(d['refs'])
        todo = [(r, level, 0) for r in reads for level in (1, 2, 3)]
        todo += [(r, level, 3) for r in reads[:4] for level in (0, 3)]   # verbosity 3: console text is part of the string
        pool = ThreadPool(8)
        res = pool.map(lambda t: mask_semi_global(ref_lib.semi_global(t[0][0], t[0][1], t[0][2], h, SCHEME, t[1], t[2])), todo)
        pool.close()
        ref_lib.delete_refs(h)
        out['sets'][setname] = [dict(read=t[0][0], sensitivity=t[1], verbosity=t[2], expected=o) for t, o in zip(todo, res)]
    write_json_gz('semiglobal_sensitivity.json.gz', out)


def make_bridge():
    """Runs tests/dropin_bridge_harness.py inside a scratch copy of the staged reference package (oracle/_ref/pydist)
    with the UNMODIFIED reference library as cpp_functions.so: chosen bridge paths + every alignment tuple."""
    import shutil
    work = '/tmp/_golden_bridge'
    shutil.rmtree(work, ignore_errors=True)
    shutil.copytree(os.path.join(ROOT, 'oracle', '_ref', 'pydist'), work)
    shutil.copy(REF_LIB, os.path.join(work, 'unicycler', 'cpp_functions.so'))
    out = os.path.join(work, 'bridge.json')
    subprocess.check_call([sys.executable, '-W', 'ignore', os.path.join(ROOT, 'tests', 'dropin_bridge_harness.py'), work, out,
                           'direct', '24'], cwd=work)
    d = json.load(open(out))
    # sequences repeat a lot (one consensus against many paths): store a string table
    table, index = [], {}

    def ref(sq):
        if sq not in index:
            index[sq] = len(table)
            table.append(sq)
        return index[sq]

    rec = [dict(fn=r['fn'], s1=ref(r['s1']), s2=ref(r['s2']), banded=r['banded'], band=r['band'], result=r['result'])
           for r in d['recorded']]
    write_json_gz('bridge_tuples.json.gz', dict(scheme=list(SCHEME), strings=table, recorded=rec, bridges=d['bridges']))


def make_dropin_align():
    """The reference's own driver (unicycler_align.semi_global_align_long_reads, 8 threads) on the sample_data fixture
    with the UNMODIFIED reference library: the alignments it keeps per read (tests/dropin_align_harness.py)."""
    import shutil
    work = '/tmp/_golden_dropin_align'
    shutil.rmtree(work, ignore_errors=True)
    shutil.copytree(os.path.join(ROOT, 'oracle', '_ref', 'pydist'), work)
    shutil.copy(REF_LIB, os.path.join(work, 'unicycler', 'cpp_functions.so'))
    out = os.path.join(work, 'align.json')
    subprocess.check_call([sys.executable, '-W', 'ignore', os.path.join(ROOT, 'tests', 'dropin_align_harness.py'), work, out, '8'],
                          cwd=work)
    d = json.load(open(out))
    write_json_gz('dropin_align_sample.json.gz', dict(reads=d['reads'], reference_seconds_8_threads=d['times'][0][1]))


def main():
    if not os.path.exists(REF_LIB):
        subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle'), 'ref'])
    ref_lib = AbiLib(REF_LIB)
    which = sys.argv[1:] or ['global', 'small', 'contained', 'tough', 'sample']
    if 'global' in which:
        make_global_path(ref_lib)
    if 'bridge' in which:
        make_bridge()
    if 'dropin_align' in which:
        make_dropin_align()
    for name, fn in (('freeend', make_freeend), ('pairs_large', make_pairs_large), ('calibration', make_calibration),
                     ('synth5', make_synth5), ('sensitivity', make_sensitivity)):
        if name in which:
            fn(ref_lib)
    sets = dict(
        small=('test/test_semi_global_alignment.fasta', 'test/test_semi_global_alignment.fastq'),
        contained=('test/test_semi_global_alignment_contained_reads.fasta',
                   'test/test_semi_global_alignment_contained_reads.fastq'),
        tough=('test/test_semi_global_alignment_tough.fasta', 'test/test_semi_global_alignment_tough.fastq'),
        sample=('sample_data/reference.fasta', 'sample_data/long_reads_low_depth.fastq.gz'))
    todo = [s for s in ('small', 'contained', 'tough', 'sample') if s in which]
    if todo:
        instr = build_instrumented()
        for s in todo:
            make_semiglobal(s, os.path.join(REF, sets[s][0]), os.path.join(REF, sets[s][1]), ref_lib, instr)


if __name__ == '__main__':
    main()
