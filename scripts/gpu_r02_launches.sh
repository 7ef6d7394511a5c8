#!/bin/bash
# ncu launch list of exactly the timed proves of the bench command (cudaProfilerStart/Stop bracket them), after a plain run
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --skip-large --skip-cpu > gpurun_out/launches_plain.json 2> gpurun_out/launches_plain.err || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_timed.csv \
    python bench.py --steps 2 --warmup 3 --skip-large --skip-cpu > gpurun_out/launches_under_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_r02_timed.csv
python - <<'P'
import json
d = json.load(open("gpurun_out/launches_plain.json"))
print("plain run:", d["value"], "ms; launches per step", d["gpu_launches"] / d["steps"], "share", d["roofline"]["share_of_step"])
P
