"""platypus_b200 — B200-native read-vs-haplotype likelihood engine (Platypus hot path)."""
from . import _abi  # noqa: F401
from .batch import Read, Window, WindowBatch, concat_batches, shard_bounds  # noqa: F401

__version__ = "0.1.0"
