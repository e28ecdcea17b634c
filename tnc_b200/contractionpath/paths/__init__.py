from .cotengrust import Cotengrust, OptMethod
from .tree_reconfiguration import TreeReconfigure, reconfigure_ssa_path, slice_and_reconfigure

__all__ = ["Cotengrust", "OptMethod", "TreeReconfigure", "reconfigure_ssa_path", "slice_and_reconfigure"]
