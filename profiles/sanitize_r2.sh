#!/bin/bash
# Round 2: the GPU parity tests under compute-sanitizer (memcheck: every test except the full-size / stress ones, which are
# the same kernels on bigger inputs; racecheck: the shared-memory heavy operators on their small cases).
# Run under gpurun; writes gpurun_out/r2_sanitizer_{memcheck,racecheck}.log.
set -u
O=gpurun_out
SKIP='not full_size and not fullsize and not stress and not config4 and not config2 and not config3 and not full_resolution and not 800_detections and not anchor_sized and not maskrcnn_tile_flow and not reference_cuda'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $O/r2_sanitizer_memcheck.log \
    python -m pytest tests -m gpu -q -x -k "$SKIP" > $O/r2_sanitizer_memcheck_pytest.log 2>&1
echo "memcheck exit $?" >> $O/r2_sanitizer_memcheck_pytest.log
tail -3 $O/r2_sanitizer_memcheck_pytest.log; tail -3 $O/r2_sanitizer_memcheck.log
RK='test_soma_binarize_vs_oracle or test_largest_cc_vs_oracle or test_paste_vs_oracle or test_nms_golden or test_peaks_golden or test_otsu_golden or test_roialign_layout or test_rle_codec'
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $O/r2_sanitizer_racecheck.log \
    python -m pytest tests -m gpu -q -x -k "$RK" > $O/r2_sanitizer_racecheck_pytest.log 2>&1
echo "racecheck exit $?" >> $O/r2_sanitizer_racecheck_pytest.log
tail -3 $O/r2_sanitizer_racecheck_pytest.log; tail -3 $O/r2_sanitizer_racecheck.log
