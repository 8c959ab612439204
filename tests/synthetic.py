"""Synthetic replay batches: re-exported from the package (deep_successor_features_for_transfer_b200/workloads.py)."""
from deep_successor_features_for_transfer_b200.workloads import synthetic_transitions  # noqa: F401
