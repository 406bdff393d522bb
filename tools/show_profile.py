import json, sys
d=json.load(open(sys.argv[1]))
tot=0
for k,v in d['segments'].items():
    tf = v['flops']/v['ms']/1e9 if v['ms']>0 else 0
    gb = v['bytes']/v['ms']/1e6 if v['ms']>0 else 0
    tot+=v['ms']
    print(f"{k:28s} kind {v['kind']} {v['ms']:8.3f} ms  {tf:8.1f} TFLOP/s  {gb:8.1f} GB/s(alg)")
print('total', tot)
