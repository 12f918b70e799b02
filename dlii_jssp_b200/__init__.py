"""dlii_jssp_b200 - B200 (sm_100a) implementation of the DLII-JSSP jerk-space sampling planner's per-cycle candidate
evaluation, behind the reference's own planner interface.

All compute goes through ``libjssp_b200.so`` (include/jssp_b200.h, built in-tree by ``dlii_jssp_b200/build.py``); there is no
CPU fallback - creating an engine without the library or a GPU raises.
"""
from . import _lib
from ._lib import JsspError, TRAJ_COLUMNS, STATE_COLUMNS
from .engine import JsspConfig, JsspEngine, CycleResult
from .frenet import State, FrenetState, JerkSpaceSamplingFrenetTrajectory, Vehicle
from .geometry import CubicSpline2D, QuinticPolynomial, spline_tables
from .planner import (JerkSpaceSamplingPlanner, JerkSpaceSamplingPlannerSettings, IntersectionPlanner, PlannerSettings)
from . import scenario
from . import distributed
from . import replay
from .fop import FrenetOptimalPlanner, FrenetOptimalPlannerSettings, GoalSamplingFrenetTrajectory
from . import batch
from .batch import BatchReplay, PackedScenario, pack_scenario, run_batch

__all__ = ["JsspConfig", "JsspEngine", "CycleResult", "JsspError", "State", "FrenetState", "JerkSpaceSamplingFrenetTrajectory", "Vehicle",
           "CubicSpline2D", "QuinticPolynomial", "spline_tables", "JerkSpaceSamplingPlanner", "JerkSpaceSamplingPlannerSettings",
           "IntersectionPlanner", "PlannerSettings", "scenario", "distributed", "replay", "FrenetOptimalPlanner", "FrenetOptimalPlannerSettings", "GoalSamplingFrenetTrajectory", "batch", "BatchReplay", "PackedScenario", "pack_scenario", "run_batch", "TRAJ_COLUMNS", "STATE_COLUMNS"]
