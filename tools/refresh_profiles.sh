#!/bin/bash
# Turn the files a GPU evidence run left in gpurun_out/ into the tracked copies under profiles/.
#   gpurun -- 'pytest -m gpu > gpurun_out/pytest_final.log; bench.py > gpurun_out/bench_r1_final.log;
#              bench.py --impl reference > gpurun_out/bench_ref_final.log;
#              ncu --metrics gpu__time_duration.sum ... --log-file gpurun_out/launches_final.csv bench.py --steps 1 --warmup 1 --no-cpu-baseline;
#              ncu --set full ... -o gpurun_out/prof_final tools/profile_step.py --rows 512 --steps 1'
set -e
cd "$(dirname "$0")/.."
R=${1:-r01}
ncu -i gpurun_out/prof_final.ncu-rep --page raw --csv > profiles/${R}_ncu_full_raw_final_512rows.csv 2>/dev/null
python tools/ncu_summary.py profiles/${R}_ncu_full_raw_final_512rows.csv > profiles/${R}_ncu_full_summary_final_512rows.csv
python tools/ncu_traffic.py profiles/${R}_ncu_full_raw_final_512rows.csv 512 > profiles/${R}_traffic.json
cp gpurun_out/launches_final.csv profiles/${R}_launches_final_bench_8192rows.csv
python tools/launch_summary.py profiles/${R}_launches_final_bench_8192rows.csv > profiles/${R}_launches_final_summary.txt
grep "^{" gpurun_out/bench_r1_final.log | python -m json.tool > profiles/${R}_bench_final.json
grep "^{" gpurun_out/bench_ref_final.log | python -m json.tool > profiles/${R}_bench_reference_arm_final.json
[ -f gpurun_out/bench_2gpu_final.log ] && grep "^{" gpurun_out/bench_2gpu_final.log | python -m json.tool > profiles/${R}_bench_2gpu_final.json
cp gpurun_out/parity_observed.json profiles/${R}_parity_observed.json
cp gpurun_out/pytest_final.log profiles/${R}_pytest_gpu_final.txt
python tools/show_bench.py gpurun_out/bench_r1_final.log
