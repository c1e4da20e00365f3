"""hopfield_network_go_b200 — B200-native hot path of hmcalister/Hopfield-Network-Go.

csrc/  hand-written sm_100a CUDA + the C ABI (include/hopfield_b200.h)
capi   ctypes binding of the C ABI (what the cgo package binds)
hopfieldnetwork  host-side mirror of the reference's Go API over the C ABI
"""
from . import build, capi  # noqa: F401
