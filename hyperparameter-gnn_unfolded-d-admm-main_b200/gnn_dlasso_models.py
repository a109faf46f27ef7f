"""Import shim: ``unfolded_train_new.py`` does ``import gnn_dlasso_models`` (reference
unfolded_train_new.py:2) without using it; the module only exists under ``old code/`` upstream."""
from gnn_dlasso_models_progressive import DLASSO_GNNHyp3_Progressive, GNNHypernetwork3  # noqa: F401
