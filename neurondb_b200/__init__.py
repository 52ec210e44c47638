"""neurondb_b200 -- B200-native vector-search hot path for NeuronDB.

The product is the C-ABI library neurondb_b200/lib/libndb_b200.so (include/ndb_b200.h);
`neurondb_b200.api` is a thin ctypes mirror of the reference's interface for that path.
"""
from . import _lib  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import (Dataset, HnswIndex, IvfIndex, NdbError, init, is_available, shutdown)  # noqa: F401

__all__ = ["Dataset", "HnswIndex", "IvfIndex", "NdbError", "init", "is_available", "shutdown"]
