import json,sys
p=json.load(open(sys.argv[1]))
print(sys.argv[1], round(p["ms_total"]/p["steps"],3), ' '.join('%s%d=%.3f'%(w[0],o["index"],o["ms"]/max(o["launches"],1)) for w in ("detector","recogniser") for o in p["ops"][w] if (w=="detector" and o["index"] in (0,2,3,8,13,24,25,26)) or (w=="recogniser" and o["index"] in (0,2,5,8,11))))
