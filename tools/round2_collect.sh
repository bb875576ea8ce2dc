#!/bin/bash
# Turn the outputs of tools/round2_profiles.sh (gpurun_out/r02_*) into the tracked evidence under profiles/.
set -e
R=gpurun_out; P=profiles
for c in config1 config2_fullsize; do
  out=$P/r02_ncu_kernels_$c.txt
  echo "ncu --set full --clock-control none --import-source on, tools/profile_step.py 500000 2 (500 k pairs = 1 M reads, second step), B200, round-2 code; $c" > $out
  n=$(ncu -i $R/r02_prof_$c.ncu-rep --page raw --csv 2>/dev/null | tail -n +3 | wc -l)
  for i in $(seq 0 $((n-1))); do echo >> $out; python tools/ncu_summary.py $R/r02_prof_$c.ncu-rep $i >> $out; done
  python tools/ncu_lanes.py $R/r02_prof_$c.ncu-rep k_search 0 40 > $P/r02_ncu_k_search_lanes_$c.txt || true
done
for c in config1 config2_fullsize config3_fullsize; do
  [ -s $R/r02_launches_$c.csv ] || continue
  cp $R/r02_launches_$c.csv $P/r02_launches_$c.csv
  python tools/launch_table.py $P/r02_launches_$c.csv 'k_encode' > $P/r02_launches_$c.txt
done
python - <<PY
import csv, json, subprocess
out = {"kernel": "k_search", "what": "DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of one k_search launch from the ncu --set full captures summarised in profiles/r02_ncu_kernels_*.txt; bench.py scales them to its launch size for roofline traffic"}
for c in ("config1", "config2_fullsize"):
    txt = subprocess.run(["ncu", "-i", f"gpurun_out/r02_prof_{c}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, u = rows[0], rows[1]
    for r in rows[2:]:
        if "k_search" in r[h.index("Kernel Name")]:
            def val(name):
                i = h.index(name); v = float(r[i].replace(",", "")); unit = u[i]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
            out[c] = {"dram_bytes_per_launch": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")), "reads_in_launch": 1000000,
                      "source": f"profiles/r02_ncu_kernels_{c}.txt"}
            break
json.dump(out, open("profiles/search_kernel_ncu.json", "w"), indent=1)
print(out)
PY
