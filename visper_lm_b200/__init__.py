"""visper_lm_b200 — B200-native (sm_100a) implementation of the VisPer-LM / OLA-VLM data-parallel
training step behind the reference's Python surface.  Kernels live in csrc/ (C ABI declared in
include/visper_b200.h); this package holds the ctypes binding, the autograd sequencing and the
host-side mirror of ola_vlm.model / ola_vlm.train for that path only."""

__version__ = "0.1.0"
