#!/bin/bash
# compute-sanitizer passes over small instances of every scatter form (memcheck: out-of-bounds / misaligned accesses,
# incl. the TMA windows of the staged kernel and the colour-major lists; racecheck on the staged kernel's shared memory)
set -u
O=gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_group.py tests/test_gpu_random_parity.py -m gpu -q -x \
  -k "ragged or c1_tridiag or kat_tridiag30 or random_patterns or group_csc_colour_shards or group_banded or invalid_and_empty or different_pattern or c4_random" \
  -p no:cacheprovider > $O/r2_sanitize_memcheck.log 2>&1
echo "memcheck exit $?" >> $O/r2_sanitize_memcheck.log
tail -6 $O/r2_sanitize_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_sanitize_racecheck.log 2>&1
echo "racecheck exit $?" >> $O/r2_sanitize_racecheck.log
tail -6 $O/r2_sanitize_racecheck.log
