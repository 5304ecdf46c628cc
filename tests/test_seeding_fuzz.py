"""Differential fuzz of the host seeding stage (line tracing with its kd-tree and point sets, seed merge, chaining)
against the reference itself: a scratch build of the reference's semi_global_align.cpp with a hook that dumps every seed
chain handed to bandedChainAlignment and then SKIPS the alignment (the DP is not what is tested here), linked against
the objects of the unmodified build (oracle/_ref).  Random references and reads of many lengths, error rates, strands,
repeats and N runs, all four sensitivity levels; the product's seed chains (ub200_seedChains, no GPU needed) must be
identical, chain by chain and seed by seed (48 cases by default; UB200_FUZZ_CASES / UB200_FUZZ_SEED for more: 90 cases
with long windows at every level, and 120 cases with seed 777, were run once with the final code — 8 and 23 minutes, no
difference).  Runs where /root/reference and the compiled reference objects exist (the
build container); skipped on the GPU box."""
import glob
import os
import random
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
REF = os.environ.get('UNICYCLER_REFERENCE', '/root/reference')
REF_SRC = os.path.join(REF, 'unicycler', 'src', 'semi_global_align.cpp')
OBJ_DIR = os.path.join(ROOT, 'oracle', '_ref', 'obj')
SCRATCH = '/tmp/unicycler_seed_fuzz'
MARKER = "        // Finally we can actually do the Seqan alignment!"
HOOK = r'''
        {
            FILE * dumpF = fopen(getenv("UNICYCLER_SEED_DUMP"), "a");
            fprintf(dumpF, "JOB\t%s%c\t%d\t%d\t%d\t%d\n", readName.c_str(), readStrand, refStart, (int)readSeq->length(), (int)trimmedRefSeq.length(), (int)length(seedChain));
            for (unsigned si = 0; si < length(seedChain); ++si)
                fprintf(dumpF, "SEED\t%ld\t%ld\t%ld\t%ld\t%ld\t%ld\n", (long)beginPositionH(seedChain[si]), (long)beginPositionV(seedChain[si]), (long)endPositionH(seedChain[si]), (long)endPositionV(seedChain[si]), (long)lowerDiagonal(seedChain[si]), (long)upperDiagonal(seedChain[si]));
            fclose(dumpF);
            continue;   // the alignment itself is not under test
        }
'''

pytestmark = pytest.mark.skipif(not (os.path.isfile(REF_SRC) and glob.glob(os.path.join(OBJ_DIR, '*.o'))),
                                reason='needs the reference sources and the compiled reference objects (build container)')


def build_hooked_reference():
    os.makedirs(SCRATCH, exist_ok=True)
    src = open(REF_SRC).read()
    assert MARKER in src
    src = src.replace(MARKER, HOOK + MARKER).replace('#include "settings.h"', '#include "settings.h"\n#include <cstdio>\n#include <cstdlib>')
    cpp, obj, lib = (os.path.join(SCRATCH, n) for n in ('semi_global_align.cpp', 'semi_global_align.o', 'libref_seedhook.so'))
    if not os.path.isfile(lib) or open(cpp).read() != src if os.path.isfile(cpp) else True:
        open(cpp, 'w').write(src)
        subprocess.check_call(['g++', '-std=c++14', '-O3', '-DNDEBUG', '-fPIC', '-w', '-I' + os.path.join(REF, 'unicycler', 'include'),
                               '-c', '-o', obj, cpp])
        objs = [o for o in glob.glob(os.path.join(OBJ_DIR, '**', '*.o'), recursive=True) if not o.endswith('/semi_global_align.o')]
        subprocess.check_call(['g++', '-shared', '-o', lib, obj] + objs + ['-lz', '-lpthread'])
    return lib


def mutate(seq, sub, ins, dele, rng):
    out = []
    for c in seq:
        r = rng.random()
        if r < sub:
            out.append(rng.choice('ACGT'))
        elif r < sub + dele:
            continue
        else:
            out.append(c)
        if rng.random() < ins:
            out.append(rng.choice('ACGT'))
    return ''.join(out)


def make_case(k, rng, level):
    # (sensitivity 2 / 3 use 9- / 8-mers: the reference's quadratic seed merge takes seconds on long windows)
    ref_len = rng.choice([1500, 4000, 9000, 20000, 45000] if level < 2 else [1500, 4000, 9000])
    ref = ''.join(rng.choice('ACGT') for _ in range(ref_len))
    if k % 5 == 1:      # a tandem repeat and a low-complexity stretch inside the reference
        unit = ''.join(rng.choice('ACGT') for _ in range(rng.choice([3, 17, 120])))
        p = rng.randrange(ref_len // 2)
        ref = ref[:p] + unit * (600 // len(unit) + 1) + 'A' * 40 + ref[p:]
    if k % 7 == 2:      # N run
        p = rng.randrange(len(ref) - 100)
        ref = ref[:p] + 'N' * rng.choice([1, 12, 60]) + ref[p:]
    L = min(len(ref) - 10, rng.choice([300, 1200, 3000, 8000, 22000] if level < 2 else [300, 1200, 3000]))
    start = rng.randrange(0, len(ref) - L)
    frag = ref[start:start + L]
    strand = rng.choice('+-')
    if strand == '-':
        from kmer_join_oracle import reverse_complement
        frag = reverse_complement(frag)
    e = rng.choice([0.02, 0.05, 0.08])
    read = mutate(frag, e, e, e, rng)
    if k % 6 == 3:      # chimera-like: a second, unrelated piece at the end
        read += ''.join(rng.choice('ACGT') for _ in range(rng.choice([200, 1500])))
    hits = '0,%d,%s,ref,%d,%d' % (len(read), strand, start, start + L)
    return ref, 'r%d' % k, read, hits, strand


def test_seed_chains_match_the_hooked_reference_on_random_inputs(ub, tmp_path):
    from refdriver import AbiLib
    from kmer_join_oracle import reverse_complement
    lib = AbiLib(build_hooked_reference())
    rng = random.Random(int(os.environ.get('UB200_FUZZ_SEED', '20240611')))
    dump = str(tmp_path / 'seeds.txt')
    os.environ['UNICYCLER_SEED_DUMP'] = dump
    checked_chains = checked_seeds = 0
    bad = []
    try:
        for k in range(int(os.environ.get('UB200_FUZZ_CASES', '48'))):
            level = k % 4
            ref, name, read, hits, strand = make_case(k, rng, level)
            if os.path.exists(dump):
                os.remove(dump)
            h = lib.new_refs([('ref', ref)])
            lib.semi_global(name, read, hits, h, (3, -6, -5, -2), sensitivity=level)
            lib.delete_refs(h)
            want, ref_start, trimmed_len = [], None, None
            if os.path.exists(dump):
                for line in open(dump):
                    p = line.rstrip('\n').split('\t')
                    if p[0] == 'JOB':
                        ref_start, trimmed_len = int(p[2]), int(p[4])
                        want.append([])
                    else:
                        want[-1].append([int(x) for x in p[1:7]])
            if ref_start is None:   # nothing reached the alignment: the product must not produce a chain either
                # the window is what getRefRange gives for a whole-read hit; recompute it like the reference does
                start, end = (int(x) for x in hits.split(',')[4:6])
                half = 1 + len(read) // 2
                ref_start, ref_end = max(0, start - half), min(len(ref), end + half)
                trimmed_len = ref_end - ref_start
            strand_seq = read if strand == '+' else reverse_complement(read)
            got = ub.seed_chains(strand_seq, ref[ref_start:ref_start + trimmed_len], level)
            # the reference stops at the first chain it rejects (empty / gap area); so does ub200_seedChains
            if got != want:
                bad.append((k, level, len(read), len(ref), len(got), len(want)))
            checked_chains += len(want)
            checked_seeds += sum(len(c) for c in want)
    finally:
        del os.environ['UNICYCLER_SEED_DUMP']
    assert not bad, bad
    assert checked_chains >= 30 and checked_seeds >= 1500, (checked_chains, checked_seeds)
