"""Grouping tail of the hot path (``src/Utils/Utils.py``: pred_to_ann :1445-1457, pred_to_person :499-514,
graph_cluster_to_persons :672-743; ``src/Utils/correlation_clustering``)."""
