"""Stand-in for torch_geometric==1.0.3 (test infrastructure; see ../README.md)."""
__version__ = "1.0.3-standin"
