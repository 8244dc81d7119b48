"""Print the per-tile clock trace written by a NA_TM_TRACE build (scripts/build_trace.sh + tc_check.py with NA_TRACE_DIR)."""
import sys, numpy as np
d = np.load(sys.argv[1]).astype(np.int64)[16:]
n_g = int(sys.argv[2]) if len(sys.argv) > 2 else 21
M = d[:352].reshape(44, 4, 2); E = d[352:352 + 44 * 12].reshape(44, 4, 3)          # (buffer of >= 896 int64, see csrc/mlp_tmem.cu)
t0 = min(x for x in list(M[:n_g].ravel()) + list(E[:n_g].ravel()) if x > 0)
f = lambda x: f'{x - t0:7d}' if x > 0 else '      -'
print('g | MMA: kb ready / issued x4 | EPI: D ready / ld done / A stored x4')
prev_e = None
for g in range(n_g):
    print(f'{g:2d} |', ' '.join(f(M[g, k, 0]) + '/' + f(M[g, k, 1]) for k in range(4)), '|', ' '.join('/'.join(f(E[g, p, w]) for w in range(3)) for p in range(4)))
print()
print('g | pass durations (D ready -> A stored)            | wait for D before pass 0 | first MMA issue after pass-0 store of previous GEMM')
for g in range(n_g):
    durs = [E[g, p, 2] - E[g, p, 0] if E[g, p, 2] > 0 and E[g, p, 0] > 0 else -1 for p in range(4)]
    ld = [E[g, p, 1] - E[g, p, 0] if E[g, p, 1] > 0 else -1 for p in range(4)]
    gap = E[g, 0, 0] - E[g - 1, 3, 2] if g > 0 and E[g - 1, 3, 2] > 0 else -1
    gap2 = M[g, 0, 0] - E[g - 1, 0, 2] if g > 0 and E[g - 1, 0, 2] > 0 else -1
    per = E[g, 0, 0] - E[g - 1, 0, 0] if g > 0 else -1
    print(f'{g:2d} | passes {durs} ld {ld} | D-wait {gap:6d} | kb0 ready after pass0 store {gap2:6d} | period {per}')
x = d[880:896]
if (x > 0).any():
    print('free-form stamps:', ' '.join(f'X{i}={int(v - t0)}' for i, v in enumerate(x) if v > 0))
