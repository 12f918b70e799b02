// jssp_kernels.cuh - the sm_100a kernels of the JSSP candidate evaluation (included by jssp_b200.cu only).
//
//   jssp_common.cuh      launch constants, BestRecord, mbarrier + TMA 1-D bulk copy helpers
//   jssp_seed.cuh        K1  seed_enum_kernel: cluster per scenario, level-by-level lists in (distributed) shared memory
//                                                           polyvt_sampling.py:167-305
//   jssp_pack.cuh        K0  pack_obstacles_kernel           obstacle dict tensors -> device tables
//   jssp_eval.cuh        K2  evaluate_group_kernel / evaluate_kernel (+ batch forms): unroll -> Frenet->Cartesian -> curvature
//                            feasibility -> rectangle collision against every obstacle mode -> cost -> argmin -> export,
//                            batched-replay hand-off         polyvt_sampling.py:311-385, jerk_space_sampling_planner.py:131-294,326-327
//   jssp_eval_small.cuh  K2b evaluate_small_kernel           one warp per candidate (default sampling grid)
//   jssp_eval_tiled.cuh  K2  evaluate_tiled_kernel (default for large cycles): two-phase tiled walk - seed-only work of a tile in
//                            parallel into the warp's shared-memory tile, then one lane per candidate
//   jssp_replay.cuh      K7  replay_persistent_kernel: one CTA per scenario loops seed enumeration -> evaluation -> selection -> export ->
//                            hand-off over the scenario's plan cycles in ONE launch        code/__main__.py:124-167, env.py:24-30
//   jssp_split.cuh           candidate-split exchange (NCCL all-gather record / peer mailboxes fused into the evaluation launch)
//   jssp_fop.cuh         K6  evaluate_fop_kernel             frenet_optimal_planner.py:97-333
//   jssp_aux.cuh         K3 dump_states_kernel, K5 project_states_kernel, K4 imm_update_kernel / predict_modes_kernel, fma_peak_kernel
#pragma once
#include <cooperative_groups.h>
#include "jssp_common.cuh"
#include "jssp_seed.cuh"
#include "jssp_pack.cuh"
#include "jssp_eval.cuh"
#include "jssp_eval_small.cuh"
#include "jssp_eval_tiled.cuh"
#include "jssp_replay.cuh"
#include "jssp_fop.cuh"
#include "jssp_aux.cuh"
