"""Summarise an `ncu --page source --csv --print-source sass` export: per kernel, top stall-sample instructions and per-opcode totals."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == 'Kernel Name':
        name = rows[i][1][:70]; hdr = rows[i + 1]; j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == 'Kernel Name'):
            if len(rows[j]) == len(hdr): body.append(rows[j])
            j += 1
        si = hdr.index('Warp Stall Sampling (All Samples)'); ai = hdr.index('Source'); ei = hdr.index('Instructions Executed')
        tot = sum(int(r[si] or 0) for r in body)
        print(f'=== {name}  instrs {len(body)}  samples {tot}')
        byop = collections.Counter(); exe = collections.Counter()
        for r in body:
            op = r[ai].split()[0] if not r[ai].strip().startswith('@') else r[ai].split()[1]
            op = op.split('.')[0]
            byop[op] += int(r[si] or 0); exe[op] += int(r[ei] or 0)
        print('  by opcode (samples%, executed warp-instr):', ', '.join(f'{k} {100*v/tot:.1f}% ({exe[k]/1e6:.1f}M)' for k, v in byop.most_common(18)))
        idx = sorted(range(len(body)), key=lambda k: -int(body[k][si] or 0))[:top]
        for k in sorted(idx):
            print(f'  [{k:5d}] {100*int(body[k][si] or 0)/tot:5.2f}%  exec {int(body[k][ei] or 0):9d}  {body[k][ai].strip()[:110]}')
        i = j
    else:
        i += 1
