"""delta_graph_slam_b200 — B200-native scan registration for delta_graph_slam's hot path.

VoxelGrid down-sampling, NDT (DIRECT1 / DIRECT7 / KDTREE) and FAST_GICP alignment and the
nearest-neighbour fitness score, behind the reference's own registration interface
(`select_registration_method`, setInputSource/Target, align, getFinalTransformation,
getFitnessScore).  The compute lives in libb200reg.so (hand-written CUDA for sm_100a behind
the C ABI in include/b200reg.h); this package is the thin host mirror used by tests and bench.
"""
from . import _lib
from ._lib import B200RegError, DIRECT1, DIRECT7, DIRECT26, KDTREE
from . import loop_batch
from .information_matrix import InformationMatrixCalculator
from .loop_detector import KeyFrame, Loop, LoopDetector, transform2Dto3D
from .odometry import FrontEnd, NativeFrontEnd, Prefilter, ScanMatchingOdometry
from .registration import DBL_MAX, DeviceCloud, FastGICP, NormalDistributionsTransform, RadiusOutlierRemoval, Registration, StatisticalOutlierRemoval, VoxelGrid, select_registration_method

__all__ = ["InformationMatrixCalculator", "KeyFrame", "Loop", "LoopDetector", "loop_batch", "transform2Dto3D", "B200RegError", "DIRECT1", "DIRECT7", "DIRECT26", "KDTREE", "DBL_MAX", "DeviceCloud", "FrontEnd", "NativeFrontEnd", "Prefilter", "ScanMatchingOdometry", "FastGICP", "NormalDistributionsTransform", "RadiusOutlierRemoval", "Registration", "StatisticalOutlierRemoval", "VoxelGrid",
           "select_registration_method"]
