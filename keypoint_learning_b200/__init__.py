"""keypoint_learning_b200 -- B200-native (sm_100a) detection path of CVLAB-Unibo/Keypoint-Learning.

Layout:
  csrc/   hand-written CUDA kernels + the C ABI of include/kpl.h   -> libkpl_b200.so
  host/   C++ facade: pcl::keypoints::KeypointLearningDetector over a PCL-free shim, TestDetector CLI
  capi.py ctypes binding + Python mirror of the detector class (what tests and bench.py drive)
  synth.py synthetic clouds of the benchmark configurations
"""
from .capi import KeypointLearningDetector, KplError, load_library, LIB_PATH  # noqa: F401

__all__ = ["KeypointLearningDetector", "KplError", "load_library", "LIB_PATH"]
