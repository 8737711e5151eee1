"""Memory instructions of an .ncu-rep source page: executed count, shared wavefronts (actual / ideal), global sectors."""
import csv, subprocess, sys
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
agg = {}
for r in rows[2:]:
    src = r[ix['Source']].strip(); parts = src.split()
    op = parts[1] if parts[0].startswith('@') else parts[0]
    if not any(op.startswith(p) for p in ('LDS', 'STS', 'ATOMS', 'LDG', 'STG', 'RED', 'ATOMG', 'LDGSTS')): continue
    f = lambda k: int(float(r[ix[k]] or 0))
    a = agg.setdefault(op, [0, 0, 0, 0, 0])
    a[0] += f('Instructions Executed'); a[1] += f('L1 Wavefronts Shared'); a[2] += f('L1 Wavefronts Shared Ideal')
    a[3] += f('L2 Theoretical Sectors Global'); a[4] += f('# Samples')
print(f"{'op':16s} {'executed':>12s} {'smem_wavefronts':>16s} {'ideal':>12s} {'glob_sectors':>14s} {'samples':>8s}")
for op, a in sorted(agg.items(), key=lambda t: -t[1][1] - t[1][3]):
    print(f"{op:16s} {a[0]:12d} {a[1]:16d} {a[2]:12d} {a[3]:14d} {a[4]:8d}")
