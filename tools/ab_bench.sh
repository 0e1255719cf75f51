#!/bin/bash
# A/B of library variants built into build/variants/*.so: resident step time and the band stages of bench.py, plus
# the tests named in $AB_TESTS (pytest -k expression) run against every variant.
for lib in default build/variants/*.so; do
  if [ "$lib" = default ]; then unset AUDIO_ANALYSIS_B200_LIB; else export AUDIO_ANALYSIS_B200_LIB=$PWD/$lib; fi
  echo "== $lib"
  if [ -n "$AB_TESTS" ]; then python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$AB_TESTS" 2>&1 | tail -1; fi
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-other-configs --no-variants 2>/dev/null | python tools/show_bench.py /dev/stdin 2>&1 | grep -E "^value|^config|band_fit|band_fused|untangle|edc " | cut -c1-180
done
