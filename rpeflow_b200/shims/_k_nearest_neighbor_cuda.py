"""Shim with the module name and symbols of the reference's pybind extension `models/csrc/_k_nearest_neighbor_cuda`
(models/csrc/k_nearest_neighbor/k_nearest_neighbor.cpp:27-29).  Drop this file into the reference's models/csrc/ (or let
rpeflow_b200.install() register it in sys.modules) and models/csrc/wrapper.py:4-8 imports the sm_100a kernels
instead of printing "Failed to load one or more CUDA extensions"."""
from rpeflow_b200.ops import _k_nearest_neighbor_cuda  # noqa: F401
