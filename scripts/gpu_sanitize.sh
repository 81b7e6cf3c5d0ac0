#!/bin/bash
# compute-sanitizer memcheck over the small-shape tests that exercise the round-2 kernels (k_conv_tma variants, k_canvas_planes,
# k_conv_rows, the cluster sampler)
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck.log \
    python -m pytest tests/test_det_tail_gpu.py tests/test_backbone_gpu.py::test_backbone_matches_golden \
    "tests/test_backbone_gpu.py::test_plane_handover_to_shrink_header_is_bit_identical" tests/test_enhancer_gpu.py -m gpu -x -q 2>&1 | tail -4
echo "rc=$?"; grep -c "Invalid\|Error" $OUT/memcheck.log; tail -5 $OUT/memcheck.log
