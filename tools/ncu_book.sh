timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02h_launches_book.csv python tools/book_once.py > /dev/null 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02h_launches_book.csv')) if len(r)>5 and r[0].isdigit()]
names=[r[4].split('(')[0].replace('<unnamed>::','').replace('void ','') for r in rows]
idx=[i for i,n in enumerate(names) if n.startswith('k_bk_keys')]
start=idx[-1]
tot=0; small=0; nsmall=0
for i in range(start, len(rows)):
    v=float(rows[i][-1].replace(',',''))/1000
    tot+=v
    if v<6: small+=v; nsmall+=1
    else: print(f"{names[i][:44]:44s} grid {rows[i][8]:>16s} {v:8.1f} us")
    if names[i].startswith('k_expand<'): break
print('launches', i-start+1, 'total', tot, 'small(<6us)', nsmall, small)
P
