"""TEST INFRASTRUCTURE — CPU restatement of the reference's common k-mer join.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this; the product never does.

Reference algorithm:
  * KmerPositions::addPositions (unicycler/src/kmers.cpp:51-65): for every start i of the read strand,
    positions[sequence.substr(i, k)].push_back(i) — an unordered_map keyed by the k-mer STRING (any character),
    per-k-mer position lists ascending.
  * alignReadToReferenceRange (unicycler/src/semi_global_align.cpp:197-207): for every start i of the trimmed
    reference window, ascending, look the window's k-mer up and push Point(readPos, i) for every readPos of its list.
  * getReverseComplement (unicycler/src/string_functions.cpp:52-79) for '-' strand ranges.

Pinned against the reference itself by tests/test_kmer_join.py: the "common k-mers: N" lines of the verbosity-3
console text the UNMODIFIED reference library printed for the golden sets (tests/golden/semiglobal_sensitivity.json.gz)
and, through the seeding as a whole, the dumped reference seed chains (tests/test_host_logic.py).
"""

_COMPLEMENT = {'A': 'T', 'T': 'A', 'G': 'C', 'C': 'G', 'R': 'Y', 'Y': 'R', 'S': 'S', 'W': 'W', 'K': 'M', 'M': 'K',
               'B': 'V', 'D': 'H', 'H': 'D', 'V': 'B', 'N': 'N', '.': '.', '-': '-', '?': '?', '*': '*'}


def reverse_complement(seq):
    """string_functions.cpp:52-79: characters outside the table are dropped."""
    return ''.join(_COMPLEMENT[c] for c in reversed(seq) if c in _COMPLEMENT)


def common_kmers(read_strand, window, k):
    """[(read position, window position), ...] in the reference's order."""
    positions = {}
    for i in range(len(read_strand) - k + 1):            # kmers.cpp:56-63
        positions.setdefault(read_strand[i:i + k], []).append(i)
    points = []
    for i in range(len(window) - k + 1):                 # semi_global_align.cpp:199-207
        for pos in positions.get(window[i:i + k], ()):
            points.append((pos, i))
    return points
