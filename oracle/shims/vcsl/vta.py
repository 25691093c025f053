"""``vcsl.vta`` surface used by vsc/baseline/localization.py:44-46,58 (oracle)."""
from oracle.tn_networkx import build_vta_model, tn  # noqa: F401
