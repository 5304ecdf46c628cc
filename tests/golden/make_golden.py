#!/usr/bin/env python3
"""Generates the golden vectors under tests/golden/ from the UNMODIFIED reference library
(oracle/_ref/libunicycler_ref.so, built from /root/reference by oracle/Makefile.ref).

Run in the build container only (it needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden.py

Outputs (all gzip-compressed JSON):
  global_path.json.gz        fullyGlobalAlignment / pathAlignment cases: the reference's own known-answer
                             inputs (test/test_cpp_wrappers.fasta, test_cpp_wrappers.py:22-125) plus seeded
                             random pairs; expected = reference output with the milliseconds field masked.
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


def main():
    if not os.path.exists(REF_LIB):
        subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle'), 'ref'])
    ref_lib = AbiLib(REF_LIB)
    which = sys.argv[1:] or ['global', 'small', 'contained', 'tough', 'sample']
    if 'global' in which:
        make_global_path(ref_lib)
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
