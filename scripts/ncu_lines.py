"""Per-CUDA-line stall samples from `ncu --page source --csv --print-source cuda,sass` (lines of the kernel's own .cu file)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want_file = sys.argv[3] if len(sys.argv) > 3 else 'mlp_tmem.cu'
cur = None; hdr = None; fpath = ''; data = collections.defaultdict(list)
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fpath = r[1]; continue
    if r[0] == 'Function Name': cur = r[1][:44]; continue
    if r[0] == 'Line No': hdr = r; continue
    if cur and hdr and len(r) == len(hdr) and r[0].isdigit() and fpath.endswith(want_file):
        data[cur].append(r)
for k, body in data.items():
    si = hdr.index('Warp Stall Sampling (All Samples)'); ei = hdr.index('Instructions Executed')
    reasons = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[si] or 0) for r in body)
    if not tot: continue
    rs = collections.Counter()
    for r in body:
        for i in reasons: rs[hdr[i]] += int(r[i] or 0)
    print('===', k, 'samples', tot, ' '.join(f'{a[6:]}:{100*b/tot:.0f}%' for a, b in rs.most_common(8)))
    best = sorted(body, key=lambda r: -int(r[si] or 0))[:top]
    for r in sorted(best, key=lambda r: int(r[0])):
        rr = sorted(((int(r[i] or 0), hdr[i][6:]) for i in reasons), reverse=True)[:2]
        print(f'{r[0]:>5} {100*int(r[si] or 0)/tot:5.2f}%  exec {int(r[ei] or 0)/1e6:8.2f}M  {rr[0][1]}/{rr[1][1]:12s} {r[1].strip()[:110]}')
