"""Import-path compatibility with reference sparsepoly/sparse_all_subsets.py."""
from .estimators import SparseAllSubsetsClassifier, SparseAllSubsetsRegressor  # noqa: F401
