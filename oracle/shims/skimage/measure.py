from oracle.leaves import moments, moments_central  # noqa: F401
