import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; idx={h:i for i,h in enumerate(hdr)}
L=[(r[idx['Kernel Name']], float(r[idx['Metric Value']].replace(',','')), r[idx['Grid Size']]) for r in rows[hi+1:] if len(r)>=len(hdr)]
adam=[i for i,(n,_,_) in enumerate(L) if 'adam_kernel' in n]
step=L[adam[1]+2:adam[2]+2]
print('launches',len(step),'total us',sum(t for _,t,_ in step)/1e3)
for n,t,g in step:
    if 'chain_tc' in n or 'wgrad_tc' in n: continue
    print(f"{t/1e3:8.1f}us {g:>14} {n[:60]}")
