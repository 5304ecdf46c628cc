"""unicycler_b200 — B200-native replacement for the long-read alignment hot path of Unicycler 0.5.1.

The product is the C-ABI shared library ``libunicycler_b200.so`` (C++ host + CUDA kernels for sm_100a,
``unicycler_b200/csrc``).  This package is the thin Python mirror of the reference's ctypes layer
(``unicycler/cpp_wrappers.py``): same function names, argument meaning and return strings, plus the
additive batch entry points.  There is no CPU fallback: importing works without a GPU (so that the host
logic can be unit-tested), but every alignment call needs a CUDA device and fails loudly otherwise.
"""
from .wrappers import (LIB_PATH, load_library, semi_global_alignment, fully_global_alignment, path_alignment,
                       get_random_sequence_alignment_mean_and_std_dev, new_ref_seqs, add_ref_seq,
                       delete_ref_seqs, fully_global_alignment_batch, path_alignment_batch,
                       chain_alignment, chain_alignment_batch, semi_global_alignment_batch, seed_chains,
                       last_stats, transfer_bytes, chain_cells, chain_plan, set_device, int_peak_ops_per_sec, ChainBench,
                       semi_global_alignment_exhaustive, start_seq_alignment, end_seq_alignment, overlap_alignment,
                       calibration_pairs, coalescer_stats, alignment_tallies, common_kmers, last_join_stats)

__all__ = ['LIB_PATH', 'load_library', 'semi_global_alignment', 'fully_global_alignment', 'path_alignment',
           'get_random_sequence_alignment_mean_and_std_dev', 'new_ref_seqs', 'add_ref_seq', 'delete_ref_seqs',
           'fully_global_alignment_batch', 'path_alignment_batch', 'chain_alignment', 'chain_alignment_batch',
           'semi_global_alignment_batch', 'seed_chains', 'last_stats', 'transfer_bytes', 'chain_cells', 'chain_plan', 'set_device', 'int_peak_ops_per_sec',
           'ChainBench', 'semi_global_alignment_exhaustive', 'start_seq_alignment', 'end_seq_alignment',
           'overlap_alignment', 'calibration_pairs', 'coalescer_stats', 'alignment_tallies', 'common_kmers',
           'last_join_stats']
