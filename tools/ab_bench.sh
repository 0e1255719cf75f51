#!/bin/bash
# A/B of library variants built into build/variants/*.so: resident step time and the band_fit stage of bench.py.
for lib in default build/variants/*.so; do
  if [ "$lib" = default ]; then unset AUDIO_ANALYSIS_B200_LIB; else export AUDIO_ANALYSIS_B200_LIB=$PWD/$lib; fi
  echo "== $lib"
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-other-configs 2>/dev/null | python tools/show_bench.py /dev/stdin 2>&1 | grep -E "^value|band_fit|untangle|edc " | cut -c1-120
done
