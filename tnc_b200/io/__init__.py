"""tnc::io -- HDF5 import / export (tnc/src/io/hdf5.rs).  The QASM front end is out of scope."""
from . import hdf5  # noqa: F401
from .hdf5 import load_data, load_tensor, store_data, store_tensor  # noqa: F401
