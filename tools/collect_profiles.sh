#!/bin/bash
# Turn the outputs of tools/final_profiles.sh (gpurun_out/final_*) into the tracked evidence under profiles/.
set -e
R=gpurun_out; P=profiles
for c in config1 config2_fullsize; do
  out=$P/r01_ncu_kernels_$c.txt
  echo "ncu --set full --clock-control none --import-source on, tools/profile_step.py 500000 2 (500 k pairs = 1 M reads, second step), B200, final round-1 code; $c" > $out
  n=$(ncu -i $R/final_prof_$c.ncu-rep --page raw --csv 2>/dev/null | tail -n +3 | wc -l)
  for i in $(seq 0 $((n-1))); do echo >> $out; python tools/ncu_summary.py $R/final_prof_$c.ncu-rep $i >> $out; done
  python tools/ncu_lanes.py $R/final_prof_$c.ncu-rep k_search 0 40 > $P/r01_ncu_k_search_lanes_$c.txt
done
for c in config1 config2_fullsize config3_fullsize; do
  cp $R/final_launches_$c.csv $P/r01_launches_$c.csv
  python tools/launch_table.py $P/r01_launches_$c.csv k_encode > $P/r01_launches_$c.txt
  cp $R/final_bench_$c.json $P/r01_bench_$c.json
done
cp $R/final_bench_reference_config1.json $P/r01_bench_reference_config1.json
[ -s $R/final_bench_config4.json ] && cp $R/final_bench_config4.json $P/r01_bench_config4.json
python - <<PY
import csv, json, subprocess
out = {"kernel": "k_search", "what": "DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of one k_search launch from the ncu --set full captures summarised in profiles/r01_ncu_kernels_*.txt; bench.py scales them to its launch size for roofline.traffic"}
for c in ("config1", "config2_fullsize"):
    txt = subprocess.run(["ncu", "-i", f"gpurun_out/final_prof_{c}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, u = rows[0], rows[1]
    for r in rows[2:]:
        if "k_search" in r[h.index("Kernel Name")]:
            def val(name):
                i = h.index(name); v = float(r[i].replace(",", "")); unit = u[i]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
            out[c] = {"dram_bytes_per_launch": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")), "reads_in_launch": 1000000,
                      "source": f"profiles/r01_ncu_kernels_{c}.txt"}
            break
json.dump(out, open("profiles/search_kernel_ncu.json", "w"), indent=1)
print(out)
PY
