import csv,sys,collections
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; im=hdr.index('Metric Name'); iv=hdr.index('Metric Value')
agg=collections.Counter()
for r in rows[1:]:
    try: agg[r[im]]+=float(r[iv].replace(',',''))
    except: pass
print({k: "%.4g"%v for k,v in agg.items()})
