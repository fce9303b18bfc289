"""Shim with the module name and symbols of the reference's pybind extension `models/csrc/_furthest_point_sampling_cuda`
(models/csrc/furthest_point_sampling/furthest_point_sampling.cpp:19-21).  Drop this file into the reference's models/csrc/ (or let
rpeflow_b200.install() register it in sys.modules) and models/csrc/wrapper.py:4-8 imports the sm_100a kernels
instead of printing "Failed to load one or more CUDA extensions"."""
from rpeflow_b200.ops import _furthest_point_sampling_cuda  # noqa: F401
