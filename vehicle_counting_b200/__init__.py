"""vehicle_counting_b200 -- B200-native (sm_100a) detect + ReID hot path for kaylode/vehicle-counting.

Layout:
  csrc/            hand-written CUDA kernels + the C ABI (libvcb200.so, include/vcb200.h)
  _lib.py, ops.py  ctypes binding of the C ABI
  engine.py        static buffer plans + CUDA-graph capture for the YOLOv5 and ReID networks
  networks/, modules/   host-side mirror of the reference call surface (drop-in for run.py)
"""
__version__ = "0.1.0"
