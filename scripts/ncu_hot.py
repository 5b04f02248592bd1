"""Top SASS instructions by stall samples / shared-memory excess wavefronts from `ncu --page source --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if len(r) > 10:
        hdr, start = r, i + 1
        break
ix = {h: k for k, h in enumerate(hdr)}
data = [r for r in rows[start:] if len(r) == len(hdr)]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, '# Samples') for r in data)
print("total samples", tot, "instrs", len(data))
print("--- top by samples")
for r in sorted(data, key=lambda r: -f(r, '# Samples'))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%6.0f %5.1f%%  exc_wf=%8.0f wf=%9.0f ideal=%9.0f  %s" % (f(r, '# Samples'), 100 * f(r, '# Samples') / tot,
          f(r, 'L1 Wavefronts Shared Excessive'), f(r, 'L1 Wavefronts Shared'), f(r, 'L1 Wavefronts Shared Ideal'), r[ix['Source']][:90]))
print("--- top by excessive shared wavefronts")
for r in sorted(data, key=lambda r: -f(r, 'L1 Wavefronts Shared Excessive'))[:12]:
    print("exc_wf=%8.0f wf=%9.0f ideal=%9.0f execs=%8.0f %s" % (f(r, 'L1 Wavefronts Shared Excessive'), f(r, 'L1 Wavefronts Shared'),
          f(r, 'L1 Wavefronts Shared Ideal'), f(r, 'Instructions Executed'), r[ix['Source']][:90]))
