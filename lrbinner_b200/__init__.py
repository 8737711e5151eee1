"""lrbinner_b200 — B200-native (sm_100a) profile stage of LRBinner: k-mer composition, the global
15-mer count table and per-read 15-mer coverage histograms, behind the reference's runner API.

    from lrbinner_b200.runners_utils import run_kmers, run_15mer_counts, run_15mer_vecs

The CUDA library (liblrb200.so) is mandatory: importing the compute modules without it raises.
"""
__version__ = "0.1.0"
