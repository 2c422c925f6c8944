"""profiles/traffic.json from an ncu --set full capture of the headline colour pass (scripts/ncu_all.sh <tag> headline):
DRAM bytes per launch of EXACTLY the module that capture ran (module_key from the capture's log line).
    python scripts/ncu_traffic.py gpurun_out/<tag>_headline.ncu-rep gpurun_out/<tag>_headline.log"""
import csv, io, json, os, re, subprocess, sys
rep, log = sys.argv[1], sys.argv[2]
line = [l for l in open(log) if "attempts/launch" in l][-1]
key = re.search(r"key (\w+)", line).group(1)
attempts = float(line.split("attempts/launch")[1])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, r = rows[0], rows[1], rows[2]
d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
def gb(k):
    v = float(d[k].replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u[k]]
rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
j = {"module_key": key, "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_written": wr,
     "algorithmic_bytes_per_launch": attempts * 36, "gpu_time_us": float(d["gpu__time_duration.sum"].replace(",", "")) * (1e3 if u["gpu__time_duration.sum"] == "ms" else 1.0),
     "source": "ncu --set full --clock-control none -k regex:mcg_pass_m1 --launch-skip 5 -c 1, scripts/ncu_cfg.py CFG=headline (8 replicas x 256^3, the launch shape of bench.py); %s" % os.path.basename(rep)}
json.dump(j, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(j))
