"""Wall time of lrb_reads_from_file on a synthetic FASTA / FASTQ vs threads (host cores of the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrbinner_b200.profile import PackedReads
rng = np.random.default_rng(1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
base = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
for fmt in ("fa", "fq"):
    path = f"{base}/lrb_ingest.{fmt}"
    with open(path, "wb") as f:
        for i in range(n if fmt == "fa" else n // 2):
            L = int(rng.gamma(2, 2500)) + 500
            s = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=L).tobytes()
            if fmt == "fa":
                f.write(b">r%d\n" % i + s + b"\n")
            else:
                f.write(b"@r%d\n" % i + s + b"\n+\n" + rng.choice(np.frombuffer(b"@>+!I5#", dtype=np.uint8), size=L).tobytes() + b"\n")
    sz = os.path.getsize(path)
    os.environ["LRB_INGEST_TRACE"] = "1"
    for t in (1, 4, 8, 16):
        t0 = time.perf_counter(); pr = PackedReads.from_file(path, threads=t); dt = time.perf_counter() - t0
        print(f"{fmt} threads={t} {dt:.3f} s {sz / dt / 1e9:.2f} GB/s reads={pr.n_reads} bases={pr.total_bases}", flush=True)
        pr.close()
    os.remove(path)
