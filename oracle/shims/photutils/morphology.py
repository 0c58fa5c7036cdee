from oracle.leaves import gini  # noqa: F401
