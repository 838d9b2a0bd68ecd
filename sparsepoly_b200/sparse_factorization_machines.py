"""Import-path compatibility with reference sparsepoly/sparse_factorization_machines.py."""
from .estimators import (  # noqa: F401
    LEARNING_RATE,
    SparseFactorizationMachineClassifier,
    SparseFactorizationMachineRegressor,
)
