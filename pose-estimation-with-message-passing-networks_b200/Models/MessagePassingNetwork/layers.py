"""Parameter containers of the message-passing layers.

These modules hold the reference's parameters under the reference's names
(``src/Models/MessagePassingNetwork/layers.py``; SURVEY.md 8b lists the
state-dict contract) so that reference checkpoints load unchanged.  They carry
no PyTorch compute: ``NodeClassificationMPNSimple.forward`` hands the packed
parameters to the CUDA path.
"""

import torch.nn as nn

NUM_TYPE_MLPS = 17  # TypeAwareNodeUpdate builds 17 MLPs whatever the dataset (layers.py:266)


def make_mlp(input_dim, hidden_dims, bn=False, end_with_relu=False):
    """Same ``nn.Sequential`` slot layout as the reference's ``_make_mlp`` (layers.py:8-29):
    Linear, then for every layer but the last ReLU [+ BatchNorm1d]; optional trailing ReLU [+ BN]."""
    mods = []
    dims = [input_dim] + list(hidden_dims)
    last = len(hidden_dims) - 1
    for i in range(len(hidden_dims)):
        mods.append(nn.Linear(dims[i], dims[i + 1]))
        if i != last:
            mods.append(nn.ReLU(inplace=True))
            if bn:
                mods.append(nn.BatchNorm1d(dims[i + 1]))
    if end_with_relu:
        mods.append(nn.ReLU(inplace=True))
        if bn:
            mods.append(nn.BatchNorm1d(dims[-1]))
    return nn.Sequential(*mods)


def _edge_mlp(node_dim, edge_dim, hidden, skip):
    f = 2 if skip else 1
    return nn.Sequential(nn.Linear(node_dim * 2 * f + edge_dim * f, hidden), nn.ReLU(inplace=True),
                         nn.Linear(hidden, edge_dim), nn.ReLU(inplace=True))


class MPLayer(nn.Module):
    """Type-agnostic layer (layers.py:32-86): ``mlp_edge``, ``mlp_node``, optional ``update_mlp``."""

    def __init__(self, node_dim, edge_dim, edge_hidden, aggr, use_node_update_mlp, skip=False, edge_mlp="agnostic"):
        super().__init__()
        if edge_mlp != "agnostic":
            raise NotImplementedError("EDGE_MLP=%r: only the agnostic edge MLP is in scope (the reference's "
                                      "per_type variants do not construct, SURVEY.md App. A)" % (edge_mlp,))
        self.aggr = aggr
        f = 2 if skip else 1
        self.mlp_edge = _edge_mlp(node_dim, edge_dim, edge_hidden, skip)
        self.mlp_node = nn.Sequential(nn.Linear(node_dim * f + edge_dim, node_dim), nn.ReLU(inplace=True))
        self.update_mlp = (nn.Sequential(nn.Linear(node_dim, node_dim), nn.ReLU())
                           if use_node_update_mlp else None)


class TypeAwareNodeUpdate(nn.Module):
    """17 message MLPs selected by the source node's type (layers.py:260-274)."""

    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.mlp = nn.ModuleList([nn.Sequential(nn.Linear(input_dim, output_dim), nn.ReLU(inplace=True))
                                  for _ in range(NUM_TYPE_MLPS)])
        self.output_dim = output_dim


class HierarchUpdateMlp(nn.Module):
    """Body-part tree over the per-type aggregates (layers.py:89-128): parameter container, reference names."""

    def __init__(self, node_dim, num_joints):
        super().__init__()
        if num_joints not in (17, 14):
            raise NotImplementedError("hierarch_mlp is defined for 17 or 14 joint types (layers.py:96)")
        first_in = 5 if num_joints == 17 else 2
        self.node_dim, self.num_joints = node_dim, num_joints
        self.first_layer = nn.ModuleList([nn.Linear(node_dim * first_in, node_dim // 2)] +
                                         [nn.Linear(node_dim * 2, node_dim // 2) for _ in range(6)])
        self.second_layer = nn.ModuleList([nn.Linear(2 * node_dim // 2, node_dim // 2) for _ in range(6)])
        self.final = nn.Linear(6 * node_dim // 2, node_dim)


class TypeAwareMPNLayer(nn.Module):
    """Per-type layer (layers.py:157-258)."""

    def __init__(self, node_dim, edge_dim, edge_hidden, aggr, skip=False, edge_mlp="agnostic", num_types=17,
                 aggr_sub=None, update_type="mlp"):
        super().__init__()
        if edge_mlp != "agnostic":
            raise NotImplementedError("EDGE_MLP=%r is out of scope" % (edge_mlp,))
        if update_type not in ("mlp", "hierarch_mlp"):
            raise NotImplementedError("UPDATE_TYPE=%r is out of scope (hierarch_cnn is a research ablation)" % (update_type,))
        if aggr_sub not in ("None", "node_edge_attn", "node_edge_attn_per_type"):
            # the reference returns None from aggregate() for anything else (layers.py:228-231)
            raise NotImplementedError("AGGR_SUB=%r" % (aggr_sub,))
        if num_types is None or num_types > NUM_TYPE_MLPS:
            raise NotImplementedError("num_types=%r" % (num_types,))
        self.aggr = aggr
        self.num_types = num_types
        self.aggr_sub = aggr_sub
        self.update_type = update_type
        f = 2 if skip else 1
        self.mlp_edge = _edge_mlp(node_dim, edge_dim, edge_hidden, skip)
        self.mlp_node = TypeAwareNodeUpdate(node_dim * f + edge_dim, node_dim)
        if update_type == "mlp":
            self.update_mlp = nn.Sequential(nn.Linear(node_dim * num_types, node_dim), nn.ReLU(inplace=True))
        else:
            self.update_mlp = HierarchUpdateMlp(node_dim, num_types)                 # layers.py:189-190
        if aggr_sub == "node_edge_attn":
            self.attn_net = nn.Sequential(nn.Linear(edge_dim, 1))
        elif aggr_sub == "node_edge_attn_per_type":
            self.attn_net = nn.Sequential(nn.Linear(edge_dim, 17))
        else:
            self.attn_net = None
