import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("N",d["n_gpus"],"value",d["value"],"ms/step",d["ms_per_step"],"x_rt",d.get("x_realtime"),"e2e ms",d["e2e"]["ms_per_step"],"e2e value",d["e2e"]["value"],"gather",d.get("final_gather_ms"))
for k in ("strong","tracking_sharded","config4_multi_gnss","multi_gnss_20msps","config5_batch_snapshots"):
    if k in d:
        v=d[k]; print(k,{kk:vv for kk,vv in v.items() if not isinstance(vv,(dict,list)) and kk not in("sharding","signals","layout","mode","code_row","metric","unit")})
