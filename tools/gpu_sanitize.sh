#!/bin/bash
# compute-sanitizer over a small slice of the GPU parity tests (run under gpurun): memcheck (out-of-bounds / misaligned
# accesses), then racecheck (shared-memory hazards in the radix, row, bin, stager, TMA-staged sweep and convex-sweep kernels).
set -u
mkdir -p gpurun_out
T="tests/test_gpu_parity.py::test_c1_stack_parity tests/test_gpu_parity.py::test_c3_terrain_mesh_parity tests/test_gpu_pair_rows.py::test_rows_around_the_short_long_threshold tests/test_gpu_sap.py::test_sap_default_world_box_and_batched_worlds tests/test_gpu_islands_deltas.py::test_deltas_and_islands_tight_mode_stack tests/test_gpu_partitioned.py::test_partitioned_bin_world_with_large_statics tests/test_gpu_partitioned.py::test_slab_ownership_and_halo_size tests/test_gpu_parity.py::test_c2_bin_parity_small tests/test_gpu_parity.py::test_c5_spheres_parity tests/test_gpu_advice.py::test_long_rows_switch_to_the_radix_passes tests/test_gpu_advice.py::test_multi_part_mesh_with_16_bit_indices tests/test_gpu_convexcast.py::test_sweeps_against_terrain_mesh_compounds_and_the_plane_branch tests/test_gpu_convexcast.py::test_sweeps_after_a_step_and_with_removed_bodies"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest $T -m gpu -x -q > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Misaligned" gpurun_out/sanitize_memcheck.log | head -20
T2="tests/test_gpu_parity.py::test_c1_stack_parity tests/test_gpu_pair_rows.py::test_rows_around_the_short_long_threshold tests/test_gpu_parity.py::test_c2_bin_parity_small tests/test_gpu_convexcast.py::test_sweeps_after_a_step_and_with_removed_bodies"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest $T2 -m gpu -x -q > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_racecheck.log | head -20
