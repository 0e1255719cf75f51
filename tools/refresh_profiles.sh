#!/bin/bash
# Turn the files a GPU evidence run left in gpurun_out/ into the tracked copies under profiles/.
#
# The evidence run (one ncu capture per gpurun call, each after the same command has exited 0 without ncu; the
# .ncu-rep stays in /tmp on the box - it is larger than gpurun_out/ may carry back - and only CSV pages come home):
#
#   gpurun -- 'python -m pytest tests -m gpu -q | tail -2 > gpurun_out/pytest_final.log;
#              python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final.log;
#              python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.log;
#              python tools/profile_step.py --rows 512 --steps 1 &&
#              ncu --set full --import-source on --clock-control none -k regex:"edc_kernel|fft_|band_|csd_|publish" -c 60 \
#                  -o /tmp/prof_final python tools/profile_step.py --rows 512 --steps 1;
#              ncu -i /tmp/prof_final.ncu-rep --page raw --csv > gpurun_out/prof_final_raw.csv;
#              ncu -i /tmp/prof_final.ncu-rep --page source --csv -k regex:band_fused > gpurun_out/fused_source.csv'
#   gpurun -- 'python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs --no-variants &&
#              ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv \
#                  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs --no-variants'
#   gpurun --gpus 2 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
#              --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_final.log'   (same for 8)
set -e
cd "$(dirname "$0")/.."
R=${1:-r02}
json() { grep "^{" "$1" | python -m json.tool > "$2"; }
cp gpurun_out/prof_final_raw.csv profiles/${R}_ncu_full_raw_final_512rows.csv
python tools/ncu_summary.py profiles/${R}_ncu_full_raw_final_512rows.csv > profiles/${R}_ncu_full_summary_final_512rows.csv
python tools/ncu_traffic.py profiles/${R}_ncu_full_raw_final_512rows.csv 512 > profiles/${R}_traffic.json
[ -f gpurun_out/fused_source.csv ] && python tools/sass_segments.py gpurun_out/fused_source.csv "25, (int)18" > profiles/${R}_band_fused_sass_segments.txt
python tools/launch_summary.py gpurun_out/launches_final.csv > profiles/${R}_launches_final_summary.txt
json gpurun_out/bench_r2_final.log profiles/${R}_bench_final.json
cp profiles/${R}_bench_final.json profiles/${R}_bench_1gpu.json
json gpurun_out/bench_ref_final.log profiles/${R}_bench_reference_arm_final.json
[ -f gpurun_out/bench_2gpu_final.log ] && json gpurun_out/bench_2gpu_final.log profiles/${R}_bench_2gpu.json
[ -f gpurun_out/bench_8gpu_final.log ] && json gpurun_out/bench_8gpu_final.log profiles/${R}_bench_8gpu.json
cp gpurun_out/parity_observed.json profiles/${R}_parity_observed.json
cp gpurun_out/pytest_final.log profiles/${R}_pytest_gpu_final.txt
python tools/show_bench.py gpurun_out/bench_r2_final.log
