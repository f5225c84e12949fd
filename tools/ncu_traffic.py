"""Join an ncu metrics CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch, captured
around tools/ncu_target.py) with gpurun_out/launch_names_b<B>.json -> per-unit DRAM traffic table.
usage: python tools/ncu_traffic.py gpurun_out/traffic.csv gpurun_out/launch_names_b<B>.json profiles/out.json"""
import csv
import json
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
names = json.load(open(sys.argv[2]))
# ncu "long" csv: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC,
# Section Name, Metric Name, Metric Unit, Metric Value
per = {}
for r in rows:
    i = int(r[0])
    per.setdefault(i, {"kernel": r[4]})[r[-3]] = (r[-2], float(r[-1].replace(",", "")))
launch = [per[i] for i in sorted(per)]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1.0, "ms": 1.0, "nsecond": 1e-6}
out = {"batch": names["batch"], "frames": names["frames"], "units": {}}
# skip anything ncu saw before the engine's first launch (torch helper kernels): align from the end
launch = launch[len(launch) - len(names["launches"]):]
for (name, alg), l in zip(names["launches"], launch):
    rd = l["dram__bytes_read.sum"]; wr = l["dram__bytes_write.sum"]; tm = l["gpu__time_duration.sum"]
    out["units"][name] = {"kernel": l["kernel"][:60], "dram_read_bytes": rd[1] * scale[rd[0]], "dram_write_bytes": wr[1] * scale[wr[0]],
                          "ncu_ms": tm[1] * scale[tm[0]], "alg_bytes": alg}
json.dump(out, open(sys.argv[3], "w"), indent=1)
tot = sum(u["dram_read_bytes"] + u["dram_write_bytes"] for u in out["units"].values())
alg = sum(u["alg_bytes"] for u in out["units"].values())
print(f"{len(out['units'])} units, dram traffic {tot / 1e9:.2f} GB vs algorithmic {alg / 1e9:.2f} GB (ratio {tot / alg:.3f})")
