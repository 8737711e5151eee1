"""Wall time of the file-level drop-ins (FASTA in, profile files out) on the GPU box, next to the reference tools on
the same file and host cores.  Usage: file_level_timing.py [n_reads] [--ref]"""
import os, shutil, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrbinner_b200 import runners_utils
from lrbinner_b200.synth import SynthSpec, write_fasta

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 200000
base = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
work = f"{base}/lrb_file_level"
shutil.rmtree(work, ignore_errors=True)
os.makedirs(work)
spec = SynthSpec(n, lengths="gamma5k", errors="ont", seed=22)
fa = f"{work}/reads.fa"
write_fasta(fa, spec.host_sequences())
sz = os.path.getsize(fa)
threads = os.cpu_count() or 8
print(f"{n} reads, {spec.total_bases / 1e6:.1f} Mbases, {sz / 1e6:.0f} MB FASTA, threads={threads}", flush=True)
os.environ["LRB_INGEST_TRACE"] = "1"
for pin in ("0", "1"):
    os.environ["LRB_PIN_READS"] = pin
    for rep in range(2):
        out = f"{work}/out_pin{pin}_{rep}"
        t0 = time.perf_counter()
        runners_utils.run_profile(fa, out, 4, 32, 10, threads, write_table=False)
        dt = time.perf_counter() - t0
        print(f"run_profile pin={pin} rep={rep}: {dt:.3f} s  {spec.total_bases / dt / 1e9:.3f} Gbases/s (file in, com_profs + cov_profs out)", flush=True)
out = f"{work}/out3"
t0 = time.perf_counter()
runners_utils.run_kmers(fa, out, 4, threads)
t1 = time.perf_counter()
runners_utils.run_15mer_counts(fa, out, threads)
t2 = time.perf_counter()
runners_utils.run_15mer_vecs(fa, out, 32, 10, threads)
t3 = time.perf_counter()
print(f"three separate drop-in calls: kmers {t1 - t0:.3f} s, 15mer counts (+4 GiB table file) {t2 - t1:.3f} s, 15mer vecs (reads table file) {t3 - t2:.3f} s", flush=True)
if "--ref" in sys.argv:
    from oracle import oracle
    if oracle.ref_available():
        ref = f"{work}/ref"
        os.makedirs(ref)
        t0 = time.perf_counter()
        oracle.ref_count_kmers(fa, f"{ref}/com_profs", 4, threads)
        t1 = time.perf_counter()
        oracle.ref_count_15mers(fa, f"{ref}/15mers-counts", threads)
        t2 = time.perf_counter()
        oracle.ref_search_15mers(f"{ref}/15mers-counts", fa, f"{ref}/cov_profs", 32, 10, threads)
        t3 = time.perf_counter()
        print(f"reference tools -t {threads}: count-kmers {t1 - t0:.3f} s, count-15mers {t2 - t1:.3f} s, search-15mers {t3 - t2:.3f} s", flush=True)
        same = all(open(f"{ref}/{f}", "rb").read() == open(f"{work}/out3/profiles/{f}", "rb").read() for f in ("com_profs", "cov_profs"))
        print("com_profs / cov_profs byte-identical to the reference tools:", same, flush=True)
shutil.rmtree(work, ignore_errors=True)
