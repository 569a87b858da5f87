"""profiles/r2_traffic.json from an ncu CSV with dram__bytes_read.sum / dram__bytes_write.sum of ipm_coop_kernel launches.

    gpurun: python scripts/dev_prof.py 4096 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
            --clock-control none -k regex:ipm_coop -c 3 --csv --log-file gpurun_out/r2/traffic.csv python scripts/dev_prof.py 4096
    here  : python scripts/ncu_traffic.py gpurun_out/r2/traffic.csv 52232 profiles/r2_traffic.json "<note>"

The launches of dev_prof.py are: root relaxation (1 child), then the 4096-child batch twice; the LAST launch is used."""
import csv, json, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr = rows[0]
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
iid = hdr.index("ID")
per = {}
for r in rows[1:]:
    if "ipm_coop" not in r[ik]:
        continue
    v = float(r[iv].replace(",", ""))
    u = r[iu].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "usecond": 1e3, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(u, 1)
    per.setdefault(int(r[iid]), {})[r[im]] = v * scale
last = per[max(per)]
iters = int(sys.argv[2])
tot = last["dram__bytes_read.sum"] + last["dram__bytes_write.sum"]
out = {"kernel": "ipm_coop_kernel", "dram_bytes_read": last["dram__bytes_read.sum"], "dram_bytes_write": last["dram__bytes_write.sum"],
       "dram_bytes_per_launch": tot, "iterations_per_launch": iters, "dram_bytes_per_iteration": tot / iters,
       "kernel_ms_under_ncu": last.get("gpu__time_duration.sum", 0) / 1e6,
       "source": (sys.argv[4] if len(sys.argv) > 4 else "") + " ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum of the 4096-child C3 launch (scripts/dev_prof.py 4096), "
                 + sys.argv[1].split("/")[-1]}
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(json.dumps(out))
