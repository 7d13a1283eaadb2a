"""Summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py: the kernels of the LAST complete
train step (between the last two Adam launches), chain / wgrad kernels first, then everything else in launch order.
usage: python tools/parse_launches.py gpurun_out/launches.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
L = [(r[idx['Kernel Name']], float(r[idx['Metric Value']].replace(',', '')), r[idx['Grid Size']]) for r in rows[hi + 1:] if len(r) >= len(hdr)]
adam = [i for i, (n, _, _) in enumerate(L) if 'adam' in n and 'p2p' not in n]
# one step = (previous Adam + increment, ..., this Adam + increment]
ends = [i for i in adam if i + 1 < len(L) and 'increment' in L[i + 1][0]]
a, b = ends[-2] + 2, ends[-1] + 2
step = L[a:b]
tot = sum(t for _, t, _ in step) / 1e3
print('launches', len(step), 'total us %.1f' % tot)
big = [(n, t, g) for n, t, g in step if 'chain_f16_kernel' in n or 'chain_tc_kernel' in n or 'wgrad_tc' in n]
print('chain + wgrad kernels: %.1f us' % (sum(t for _, t, _ in big) / 1e3))
for n, t, g in big:
    print(f"{t/1e3:8.1f}us {g:>14} {n[:70]}")
rest = [(n, t, g) for n, t, g in step if not ('chain_f16_kernel' in n or 'chain_tc_kernel' in n or 'wgrad_tc' in n)]
print('everything else (stem, transitions, head, packs, folds, reductions, amax, Adam): %.1f us in %d launches' % (sum(t for _, t, _ in rest) / 1e3, len(rest)))
for n, t, g in rest:
    print(f"{t/1e3:8.1f}us {g:>14} {n[:70]}")
