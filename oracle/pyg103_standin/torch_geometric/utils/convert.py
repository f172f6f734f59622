"""`from torch_geometric.utils import convert` (chem/util.py:6) — imported at module top, never used on the path."""
