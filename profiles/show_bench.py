import json,sys
for f in sys.argv[1:]:
    r=json.loads(open(f).read().strip().split("\n")[-1])
    def show(k,x): print(f,k,"value",round(x["value"]),"ms",round(x["ms_per_step"],4),"frac",round(x["roofline"]["frac"],3),{a:round(b*1e3,1) for a,b in x["roofline"]["stage_ms"].items()},"e2e ms",round(x["e2e"]["ms_per_step"],4),"p99",round(x["latency_ms"]["p99"],4), x.get("clocks",{}).get("sm_mhz"))
    show("c2",r); [show(k,v) for k,v in r.get("also",{}).items()]
