"""Drop-in for the reference's bio/model.py (same class names, constructors, forward signatures, state_dict keys)."""
from .model import GNN, GNN_graphpred, GINConv, GCNConv, GATConv, GraphSAGEConv, global_mean_pool  # noqa: F401
