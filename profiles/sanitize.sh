#!/bin/sh
# compute-sanitizer over the parity tests that exercise every kernel on small scenes (run on the GPU box from the repository
# root; logs land in gpurun_out/ and are committed under profiles/).  memcheck: out-of-bounds / misaligned accesses and API
# errors; racecheck: shared-memory hazards (the tiled sweeps' TMA-staged images, the sort's staging area, the plan's tables).
set -x
SEL_MEM='test_edge_scene_whole_steps or test_smallest_handle or test_golden_vectors_cuda or test_golden_edge_vectors_cuda or test_cuda_against_reference_golden or test_virtual_slabs_device_side_counts or test_virtual_slabs_rebalancing or test_virtual_slabs_canonical_order_is_bit_exact or test_canonical_order_is_path_independent or test_option_full_support_search or test_stage_api_out_of_order or test_sort_pairs_stable or test_resume_is_bit_exact'
SEL_RACE='(test_virtual_slabs_canonical_order_is_bit_exact and 2) or test_smallest_handle or test_golden_vectors_cuda or test_golden_edge_vectors_cuda or (test_virtual_slabs_device_side_counts and graph) or (test_sort_pairs_stable and 4097)'
compute-sanitizer --tool memcheck --error-exitcode 77 --print-limit 20 python -m pytest tests -m gpu -q -x -k "$SEL_MEM" > gpurun_out/${1:-r02}_memcheck.log 2>&1
echo "memcheck exit code: $?" >> gpurun_out/${1:-r02}_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 77 --print-limit 20 python -m pytest tests -m gpu -q -x -k "$SEL_RACE" > gpurun_out/${1:-r02}_racecheck.log 2>&1
echo "racecheck exit code: $?" >> gpurun_out/${1:-r02}_racecheck.log
tail -n 4 gpurun_out/${1:-r02}_memcheck.log; tail -n 4 gpurun_out/${1:-r02}_racecheck.log
