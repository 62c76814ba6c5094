"""qcgraft -- B200-native inspection hot path (K1 preprocess, K2 YOLOv8 decode+NMS, K3 ROI-align,
K4 defect masking) behind the reference's Python call signatures.

Sub-modules (import as ``qcgraft.<name>``):
  build     nvcc build of csrc/ -> lib/libqcgraft.so (sm_100a)
  ffi       ctypes binding of include/qcgraft.h + tensor-level wrappers
  nets      plain-torch YOLOv8 / ResNet-50 graphs (random init; outside the graft target)
  synth     seeded synthetic inputs for the five BASELINE configs
  pipeline  batched device pipeline K1 -> net -> K2 -> K3 -> net -> K4
  detector  drop-in classes with the reference's signatures and result dicts
  dist      one-process-per-GPU sharding + the SPC all-reduce
There is no CPU fallback: anything that computes raises if libqcgraft.so cannot be loaded.
"""
__version__ = "0.1.0"
