"""Aggregate an .ncu-rep source page by instruction opcode and show the executed-instruction mix (dynamic)."""
import csv, subprocess, sys, collections
def main(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    ops = collections.Counter(); samp = collections.Counter()
    tot = 0
    for r in rows[2:]:
        src = r[ix['Source']].strip()
        parts = src.split()
        op = parts[1] if parts[0].startswith('@') else parts[0]
        op = op.split('.')[0]
        e = int(r[ix['Instructions Executed']] or 0)
        ops[op] += e; tot += e; samp[op] += int(r[ix['# Samples']] or 0)
    st = sum(samp.values())
    for op, e in ops.most_common(25):
        print(f"{op:10s} {e:>12d} {100*e/tot:5.1f}%  samples {100*samp[op]/st:5.1f}%")
if __name__ == '__main__':
    main(sys.argv[1])
