"""Per-instruction view of an .ncu-rep source page: address, SASS, samples, executed, top stall — hot rows only."""
import csv, subprocess, sys
def main(rep, top=40):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith('stall_') and '(Not Issued)' not in h]
    body = rows[2:]
    tot = sum(int(r[ix['# Samples']] or 0) for r in body)
    texec = sum(int(r[ix['Instructions Executed']] or 0) for r in body)
    print(f"total samples {tot}, warp instructions {texec}, static instructions {len(body)}")
    ranked = sorted(enumerate(body), key=lambda t: -int(t[1][ix['# Samples']] or 0))[:int(top)]
    for n, r in sorted(ranked):
        s = int(r[ix['# Samples']] or 0)
        st = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
        print(f"{n:5d} {100*s/max(tot,1):5.1f}% exec={r[ix['Instructions Executed']]:>9} {r[ix['Source']].strip()[:70]:70s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")
if __name__ == '__main__':
    main(*sys.argv[1:])
