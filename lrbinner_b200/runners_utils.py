"""Drop-in for the runner functions of mbcclr_utils/runners_utils.py (reference lines 78-113).

Same names, same arguments, same side effects: each call makes `{output}/profiles` if needed, blocks
until its file is complete, returns None, and on failure goes through check_proc -> logger.error x2
+ sys.exit(ret).  Instead of os.system() on the three C++/OpenMP executables the work is done by
liblrb200.so (CUDA, sm_100a) through ctypes:

    run_kmers        -> lrb_count_kmers     (count-kmers    <reads> <out> <k> <threads>)
    run_15mer_counts -> lrb_count_15mers    (count-15mers   <reads> <out> <threads>)
    run_15mer_vecs   -> lrb_search_15mers   (search-15mers  <table> <reads> <out> <bin_size> <bins> <threads>)

`threads` only sizes the host-side parser/packer/formatter pools; the profile arithmetic runs on the GPU.
A maintainer switches LRBinner over with one import line in mbcclr_utils/pipelines.py.
run_profile() is the fused fast path for stages 1_1 + 1_2 + 2_1 (parse once, table stays in HBM).
"""
import logging
import os
import sys

from . import _lib

logger = logging.getLogger('LRBinner')


def run_kmers(reads_path, output, k_size, threads):
    if not os.path.isdir(f"{output}/profiles"):
        os.makedirs(f"{output}/profiles")

    logger.debug(f"LIB::lrb_count_kmers \"{reads_path}\" \"{output}/profiles/com_profs\" {k_size} {threads}")
    o = _lib.lib.lrb_count_kmers(os.fsencode(reads_path), os.fsencode(f"{output}/profiles/com_profs"), int(k_size), int(threads))
    check_proc(o, "Counting Trimers")


def run_15mer_counts(reads_path, output, threads):
    if not os.path.isdir(f"{output}/profiles"):
        os.makedirs(f"{output}/profiles")

    logger.debug(f"LIB::lrb_count_15mers \"{reads_path}\" \"{output}/profiles/15mers-counts\" {threads}")
    o = _lib.lib.lrb_count_15mers(os.fsencode(reads_path), os.fsencode(f"{output}/profiles/15mers-counts"), int(threads))
    check_proc(o, "Counting 15-mers")


def run_15mer_vecs(reads_path, output, bin_size, bin_count, threads):
    if not os.path.isdir(f"{output}/profiles"):
        os.makedirs(f"{output}/profiles")

    logger.debug(f"LIB::lrb_search_15mers \"{output}/profiles/15mers-counts\" \"{reads_path}\" \"{output}/profiles/cov_profs\" {bin_size} {bin_count} {threads}")
    o = _lib.lib.lrb_search_15mers(os.fsencode(f"{output}/profiles/15mers-counts"), os.fsencode(reads_path),
                                   os.fsencode(f"{output}/profiles/cov_profs"), int(bin_size), int(bin_count), int(threads))
    check_proc(o, "Counting 15-mer profiles")


def run_profile(reads_path, output, k_size, bin_size, bin_count, threads, write_table=True, write_npy=False, n_gpus=0):
    """Fused stages 1_1 + 1_2 + 2_1: leaves com_profs, cov_profs (and 15mers-counts unless write_table=False,
    needed by --resume with changed -bs/-bc) in {output}/profiles, byte-identical to the three separate calls.
    n_gpus > 1 shards the reads over that many GPUs of this node (0: LRB_GPUS from the environment, default 1);
    the three separate runners above honour LRB_GPUS the same way.  Same files whatever the GPU count."""
    if not os.path.isdir(f"{output}/profiles"):
        os.makedirs(f"{output}/profiles")

    logger.debug(f"LIB::lrb_profile \"{reads_path}\" \"{output}\" {k_size} {bin_size} {bin_count} {threads}")
    o = _lib.lib.lrb_profile_multi(os.fsencode(reads_path), os.fsencode(output), int(k_size), int(bin_size), int(bin_count),
                                   int(threads), int(n_gpus), 1 if write_table else 0, 1 if write_npy else 0)
    check_proc(o, "Computing profiles")


def check_proc(ret, name=""):
    if ret != 0:
        if name != "":
            logger.error(f"Error in step: {name}")
        logger.error(f"Failed due to an error. Please check the log. Good Bye! ({_lib.last_error()})")
        sys.exit(ret)
