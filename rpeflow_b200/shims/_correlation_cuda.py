"""Shim with the module name and symbols of the reference's pybind extension `models/csrc/_correlation_cuda`
(models/csrc/correlation/correlation.cpp:38-41).  Drop this file into the reference's models/csrc/ (or let
rpeflow_b200.install() register it in sys.modules) and models/csrc/wrapper.py:4-8 imports the sm_100a kernels
instead of printing "Failed to load one or more CUDA extensions"."""
from rpeflow_b200.ops import _correlation_forward_cuda, _correlation_backward_cuda  # noqa: F401
