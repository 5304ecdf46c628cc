"""CPU tests of the product's host side: the C-ABI library loads and exports every declared symbol, the
seeding stage reproduces the reference's seed chains, the chain planner and the storage-geometry helpers
agree with the oracle's literal emulation.  No DP call is made (that needs a GPU)."""
import ctypes
import os
import re
import subprocess

import pytest

from oracle_lib import ROOT, golden_chain_jobs, load_golden


def test_library_exports_every_declared_symbol(ub):
    header = open(os.path.join(ROOT, 'include', 'unicycler_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    names = set(re.findall(r'\b([A-Za-z_][A-Za-z0-9_]*)\s*\(', header)) - {'defined'}
    assert {'semiGlobalAlignment', 'fullyGlobalAlignment', 'pathAlignment', 'getRandomSequenceAlignmentScores',
            'newRefSeqs', 'addRefSeq', 'deleteRefSeqs', 'freeCString'} <= names
    lib = ctypes.CDLL(ub.LIB_PATH)
    for n in sorted(names):
        assert hasattr(lib, n), n
    # every symbol unicycler/cpp_wrappers.py dereferences at import time must resolve too
    for n in ['semiGlobalAlignmentExhaustive', 'startAlignment', 'endAlignment', 'overlapAlignment',
              'multipleSequenceAlignment', 'minimapAlignReads', 'minimapAlignReadsWithSettings', 'miniasmAssembly',
              'simulateDepths', 'getRandomSequenceAlignmentErrorRates']:
        assert hasattr(lib, n), n


def test_ref_seq_handles(ub):
    h = ub.new_ref_seqs()
    ub.add_ref_seq(h, 'a', 'ACGT')
    ub.add_ref_seq(h, 'b', 'TTTT')
    ub.delete_ref_seqs(h)


@pytest.mark.parametrize('setname', ['small', 'contained', 'tough'])
def test_seeding_matches_reference_seed_chains(ub, setname):
    d = load_golden('semiglobal_%s.json.gz' % setname)
    groups, order = {}, []
    for j in golden_chain_jobs(d):
        k = (j['readName'], j['refName'], j['refOffset'])
        if k not in groups:
            groups[k] = dict(read=j['readSeq'], ref=j['refSeq'], chains=[])
            order.append(k)
        groups[k]['chains'].append(j['seeds'])
    assert order
    for k in order:
        g = groups[k]
        assert ub.seed_chains(g['read'], g['ref'], 0) == g['chains'], k


def _build_cpp(name, sources):
    out = os.path.join('/tmp', name)
    subprocess.check_call(['g++', '-std=c++17', '-O2', '-I/usr/local/cuda/include', '-o', out] +
                          [s if s.startswith('-') else os.path.join(ROOT, s) for s in sources])
    return out


def test_host_thread_pool():
    exe = _build_cpp('ub200_test_hostpool', ['tests/cpp/test_hostpool.cpp', 'unicycler_b200/csrc/hostpool.cpp', '-lpthread'])
    for threads in ('1', '3', '16'):
        out = subprocess.run([exe], env=dict(os.environ, UNICYCLER_B200_HOST_THREADS=threads), stdout=subprocess.PIPE,
                             timeout=120)
        assert out.returncode == 0 and b' bad 0' in out.stdout, out.stdout


def test_storage_geometry_matches_oracle_navigator():
    exe = _build_cpp('ub200_test_geom', ['tests/cpp/test_geom.cpp', 'oracle/dp_oracle.cpp'])
    out = subprocess.check_output([exe, '8000']).decode()
    assert ' bad 0' in out, out


def test_point_set_iterates_like_the_reference_container():
    """pointset.hpp against std::unordered_set<Point, PointHash>: same elements in the same order after range
    construction, range / single insertions with duplicates, and copies (doubles are summed in that order)."""
    exe = _build_cpp('ub200_test_pointset', ['tests/cpp/test_pointset.cpp'])
    out = subprocess.check_output([exe, '300']).decode()
    assert ' bad 0' in out, out


@pytest.mark.parametrize('setname', ['small', 'contained', 'tough'])
def test_chain_planner_matches_oracle_grid_sequence(setname):
    exe = _build_cpp('ub200_test_plan', ['tests/cpp/test_plan.cpp', 'oracle/dp_oracle.cpp',
                                         'unicycler_b200/csrc/host_align.cpp'])
    d = load_golden('semiglobal_%s.json.gz' % setname)
    jobs = d['jobs'] if setname != 'tough' else [j for j in d['jobs'] if j['readLen'] < 9000][:12]
    lines = []
    for j in jobs:
        lines.append('JOB %d %d %d %d' % (j['readLen'], j['refLen'], j['band'], len(j['seeds'])))
        lines += [' '.join(str(x) for x in s) for s in j['seeds']]
    out = subprocess.run([exe], input='\n'.join(lines).encode(), stdout=subprocess.PIPE, check=True).stdout.decode()
    assert ' bad 0' in out and 'jobs %d ' % len(jobs) in out, out


def test_chain_plan_is_consistent_with_cell_count(ub):
    """ub200_chainPlan (planner introspection) lists the same sub-DPs whose reference cell counts ub200_chainCells sums."""
    d = load_golden('semiglobal_small.json.gz')
    for j in golden_chain_jobs(d)[:5]:
        plan = ub.chain_plan(len(j['readSeq']), len(j['refSeq']), j['seeds'], j['band'])
        cells, n = ub.chain_cells(len(j['readSeq']), len(j['refSeq']), j['seeds'], j['band'])
        assert len(plan) == n
        total = 0
        for kind, nH, nV, banded, lo, up, h0, v0, hNext, vNext in plan:
            assert 0 <= h0 and h0 + nH <= len(j['readSeq']) and 0 <= v0 and v0 + nV <= len(j['refSeq'])
            if banded:
                dimv = min(nV + 1, min(nH, up) - max(lo, -nV) + 1)
                total += (nH + 1 - max(0, lo)) * dimv
            else:
                total += (nH + 1) * (nV + 1)
        assert total == cells


def test_gap_area_guard_in_host_seeding(ub):
    """A seed chain whose largest gap exceeds MAX_BANDED_ALIGNMENT_GAP_AREA (1e8 cells) ends the range without an
    alignment (semi_global_align.cpp:286-291): two 1.5 kb matching blocks 10.5 kb of unrelated sequence apart."""
    import random
    rng = random.Random(5)

    def rs(n):
        return ''.join(rng.choice('ACGT') for _ in range(n))

    sizes = {}
    for gap in (9000, 10500):
        a, b = rs(1500), rs(1500)
        read = a + rs(gap) + b
        ref = rs(300) + a + rs(gap) + b + rs(300)
        sizes[gap] = len(ub.seed_chains(read, ref, 0))
    assert sizes[9000] == 1 and sizes[10500] == 0, sizes
