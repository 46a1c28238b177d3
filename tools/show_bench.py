import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f)); b=d["roofline"]["breakdown_ms"]
        print(f.split('/')[-1], round(d["value"]), "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in b.items()}, "e2e", round(d["e2e"]["value"]))
    except Exception as e: print(f, "ERR", e)
