"""Config objects carrying the attribute names the reference reads.

The reference uses yacs ``CfgNode`` trees (``src/config/default_config.py:116-168``);
yacs is not a dependency here.  Anything with the same attributes works (a yacs
node, a ``SimpleNamespace``, ...); these helpers build ``SimpleNamespace`` trees
with the reference's defaults so that tests / bench / users do not need yacs.
"""

from types import SimpleNamespace as NS


def default_gc_config(**overrides):
    """``_C.MODEL.GC`` defaults, default_config.py:147-168."""
    cfg = NS(
        NAME="NaiveGraphConstructor", POOL_KERNEL_SIZE=3, CHEAT=False, USE_GT=False, USE_NEIGHBOURS=False,
        EDGE_LABEL_METHOD=4, MASK_CROWDS=True, DETECT_THRESHOLD=0.005, WITH_BACKGROUND=False, HYBRID_K=5,
        MATCHING_RADIUS=0.1, INCLUSION_RADIUS=0.75, GRAPH_TYPE="knn", CC_METHOD="GAEC",
        NORM_NODE_DISTANCE=False, IMAGE_CENTRIC_SAMPLING=False, NODE_MATCHING_RADIUS=0.5,
        NODE_INCLUSION_RADIUS=0.7, WEIGHT_CLASS_LOSS=False,
        EDGE_FEATURES_TO_USE=["position", "connection_type"], NODE_DROPOUT=0.0,
    )
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


def bench_gc_config(k=30, graph_type="knn", **overrides):
    """GC settings that yield exactly ``k`` candidates per joint through the
    reference code."""
    base = dict(DETECT_THRESHOLD=1.0, HYBRID_K=k, POOL_KERNEL_SIZE=5, MASK_CROWDS=False,
                NORM_NODE_DISTANCE=True, GRAPH_TYPE=graph_type)
    base.update(overrides)
    return default_gc_config(**base)


def default_mpn_config(num_joints=17, **overrides):
    """``_C.MODEL.MPN`` defaults (default_config.py:116-142) completed with the
    sub-nodes every ``NodeClassificationMPN`` YAML supplies (e.g.
    experiments/hybrid_class_agnostic_end2end/model_58_4.yaml:91-137)."""
    cfg = NS(
        NAME="NodeClassificationMPN", NODE_TYPE_SUMMARY="not", STEPS=10, NODE_STEPS=0, EDGE_MLP="agnostic",
        NODE_INPUT_DIM=128, AGGR_TYPE="agnostic", EDGE_INPUT_DIM=num_joints + 2, EDGE_FEATURE_DIM=64,
        EDGE_FEATURE_HIDDEN=64, NODE_FEATURE_DIM=64, USE_NODE_UPDATE_MLP=False, BN=True, AGGR="max",
        AGGR_SUB="None", UPDATE_TYPE="mlp", SKIP=False, AUX_LOSS_STEPS=0, DROP_FEATURE="", EDGE_STEPS=0,
        LATE_FUSION_POS=False, NUM_JOINTS=num_joints, NODE_THRESHOLD=0.1,
        NODE_EMB=NS(BN=True, END_WITH_RELU=False, OUTPUT_SIZES=[128, 64, 64]),
        EDGE_EMB=NS(BN=True, END_WITH_RELU=False, OUTPUT_SIZES=[32, 64, 64, 64]),
        EDGE_CLASS=NS(BN=True, OUTPUT_SIZES=[64, 32, 1]),
        NODE_CLASS=NS(BN=True, OUTPUT_SIZES=[64, 32, 1]),
        CLASS=NS(BN=True, OUTPUT_SIZES=[64, 32, num_joints]),
    )
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


def flagship_mpn_config(num_joints=17, **overrides):
    """hybrid_class_agnostic_end2end/model_58_4.yaml:91-137: per-type messages,
    edge attention, skip connections, 10 steps."""
    base = dict(AGGR_TYPE="per_type", AGGR="add", AGGR_SUB="node_edge_attn", SKIP=True, STEPS=10, BN=False)   # BN: False, model_58_4.yaml:135
    base.update(overrides)
    return default_mpn_config(num_joints, **base)


def agnostic_mpn_config(num_joints=17, **overrides):
    """class_agnostic_end2end/model_57_1_0.yaml shape: agnostic MPLayer, max aggregation, skip."""
    base = dict(AGGR_TYPE="agnostic", AGGR="max", SKIP=True, STEPS=10, BN=False)   # BN: False in the YAML (the code default is True, default_config.py:132)
    base.update(overrides)
    return default_mpn_config(num_joints, **base)
