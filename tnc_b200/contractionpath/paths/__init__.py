from .cotengrust import Cotengrust, OptMethod

__all__ = ["Cotengrust", "OptMethod"]
