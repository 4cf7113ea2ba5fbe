"""fluidity_b200: B200-native CG element assembly behind Fluidity's element-loop seam.

Only what the hot path needs lives here: csrc/ (sm_100a CUDA kernels + the C ABI of
include/cgasm.h), the ctypes host binding (cgasm.py), and synthetic meshes/fields for the
parity tests and bench (synthetic.py). See DESIGN.md.
"""
