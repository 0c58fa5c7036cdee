from oracle.leaves import resize, rotate  # noqa: F401
