"""Drop-in for ``Models.MessagePassingNetwork`` of the reference
(``src/Models/MessagePassingNetwork/__init__.py:27-73``): ``get_mpn_model(config)``.

``NAME == "NodeClassificationMPN"`` (81 of the reference's 227 configs, both published
checkpoints) is served by the CUDA path; the other 18 research variants are out of scope and
raise ``NotImplementedError`` -- there is no silent PyTorch fallback.
"""

import ctypes as C

import torch
import torch.nn as nn

from ... import _native as nv
from .layers import MPLayer, TypeAwareMPNLayer, make_mlp

_NODE_TYPE_MAPS = {  # MessagePassingNetwork/utils.py:6-19
    "left_right": [0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8],
    "per_body_part": [0, 0, 0, 0, 0, 1, 1, 2, 3, 2, 3, 4, 5, 4, 5, 4, 5],
}


def sum_node_types(node_summary, node_types):
    if node_summary == "not":
        return node_types
    if node_summary in _NODE_TYPE_MAPS:
        return torch.tensor(_NODE_TYPE_MAPS[node_summary], device=node_types.device)[node_types]
    raise NotImplementedError(node_summary)


class _Packer:
    """Lays tensors out in one flat fp32 device buffer (256-byte aligned pieces)."""

    def __init__(self):
        self.pieces, self.offset = [], 0

    def add(self, t):
        t = t.detach().to(torch.float32).contiguous().reshape(-1)
        off = self.offset
        self.pieces.append((off, t))
        self.offset = (off + t.numel() + 63) // 64 * 64
        return off

    def finish(self, device):
        flat = torch.zeros(max(self.offset, 64), dtype=torch.float32, device=device)
        for off, t in self.pieces:
            flat[off:off + t.numel()] = t.to(device)
        return flat


def _fold_mlp(seq, packer):
    """``make_mlp`` chain -> list of (offset_Wt, offset_b, K, O, relu) with eval-mode BatchNorm folded into
    the following Linear (BN(v) = s*v + t  =>  W(s*v + t) + b = (W diag s) v + (W t + b))."""
    layers, pending, post = [], None, None
    mods = list(seq)
    for i, m in enumerate(mods):
        if isinstance(m, nn.Linear):
            W, b = m.weight.detach().float(), m.bias.detach().float()
            if pending is not None:
                s, t = pending
                b = b + W @ t
                W = W * s[None, :]
                pending = None
            layers.append([W, b, 0])
        elif isinstance(m, nn.ReLU):
            layers[-1][2] = 1
        elif isinstance(m, nn.BatchNorm1d):
            s = m.weight.detach().float() / torch.sqrt(m.running_var.detach().float() + m.eps)
            t = m.bias.detach().float() - m.running_mean.detach().float() * s
            if any(isinstance(n, nn.Linear) for n in mods[i + 1:]):
                pending = (s, t)
            else:
                post = (s, t)
        else:
            raise NotImplementedError(type(m))
    out = [(packer.add(W.t()), packer.add(b), W.shape[1], W.shape[0], r) for W, b, r in layers]
    post_off = (packer.add(post[0]), packer.add(post[1])) if post is not None else None
    return out, post_off, [W for W, _, _ in layers]


def _mlp_struct(spec, base):
    layers, post = spec[0], spec[1]
    if len(layers) > nv.MAX_LAYERS:
        raise NotImplementedError("MLP deeper than %d layers" % nv.MAX_LAYERS)
    m = nv.Mlp()
    m.n_layers = len(layers)
    m.dims[0] = layers[0][2]
    for l, (ow, ob, K, O, r) in enumerate(layers):
        m.dims[l + 1], m.relu[l] = O, r
        m.wt[l], m.bias[l] = base + 4 * ow, base + 4 * ob
    if post is not None:
        m.post_scale, m.post_shift = base + 4 * post[0], base + 4 * post[1]
    return m


def _mlp_train_struct(seq, prefix, offsets):
    """``make_mlp`` chain -> ``pgmp_mlp_train`` (training mode: BatchNorm stays a separate stage, nothing is folded)."""
    m = nv.MlpTrain()
    l = -1
    for i, mod in enumerate(seq):
        name = "%s.%d" % (prefix, i)
        if isinstance(mod, nn.Linear):
            l += 1
            if l >= nv.MAX_LAYERS:
                raise NotImplementedError("MLP deeper than %d layers" % nv.MAX_LAYERS)
            m.dims[l], m.dims[l + 1] = mod.in_features, mod.out_features
            m.w[l], m.b[l] = offsets[name + ".weight"], offsets[name + ".bias"]
        elif isinstance(mod, nn.ReLU):
            m.relu[l] = 1
        elif isinstance(mod, nn.BatchNorm1d):
            if mod.momentum != 0.1 or mod.eps != 1e-5 or not mod.affine or not mod.track_running_stats:
                raise NotImplementedError("BatchNorm1d with non-default settings")
            m.bn[l] = 1
            m.gamma[l], m.beta[l] = offsets[name + ".weight"], offsets[name + ".bias"]
            m.running_mean[l], m.running_var[l] = mod.running_mean.data_ptr(), mod.running_var.data_ptr()
        else:
            raise NotImplementedError(type(mod))
    m.n_layers = l + 1
    return m


class _MpnTrainFunction(torch.autograd.Function):
    """One training step of the MPN in ``libpgmp.so``: ``pgmp_mpn_train_forward`` / ``pgmp_mpn_train_backward``
    stand where torch autograd runs the reference's Python forward (train.py:232-236)."""

    @staticmethod
    def forward(ctx, model, x, edge_attr, edge_index, node_types, *params):
        names = [n for n, _ in model.named_parameters()]
        dev = x.device
        lib = nv.lib()
        sizes = [p.numel() for p in params]
        offsets, off, pieces = {}, 0, []
        pad = torch.zeros(3, dtype=torch.float32, device=dev)
        for n, k, q in zip(names, sizes, params):      # every tensor starts on a 16-byte boundary (vectorised loads)
            offsets[n] = off
            pieces.append(q.detach().reshape(-1))
            if k % 4:
                pieces.append(pad[:4 - k % 4])
            off += (k + 3) // 4 * 4
        flat = torch.cat(pieces).contiguous()
        x_ = x.detach().contiguous()
        ea = edge_attr.detach().contiguous()
        ei = edge_index.detach().contiguous()
        N, E = x_.shape[0], ei.shape[1]
        n_out = model.num_outputs()
        J = model.classification[-1].out_features
        edge_logits = torch.empty((n_out, E), dtype=torch.float32, device=dev)
        node_logits = torch.empty((n_out, N), dtype=torch.float32, device=dev)
        class_logits = torch.empty((n_out, N, J), dtype=torch.float32, device=dev)
        layer = model.mpn_node_cls
        p = nv.MpnTrainParams(num_nodes=N, num_edges=E, x=x_.data_ptr(), edge_attr=ea.data_ptr(), edge_index=ei.data_ptr(),
                              dim=64, skip=int(bool(model.use_skip_connections)), steps=model.edge_steps,
                              aux_loss_steps=model.aux_loss_steps, aggr=nv.AGGR[model.aggr],
                              has_update_mlp=int(layer.update_mlp is not None), num_classes=J, params=flat.data_ptr(),
                              edge_logits=edge_logits.data_ptr(), node_logits=node_logits.data_ptr(),
                              class_logits=class_logits.data_ptr())
        for field, name in (("node_emb", "node_embedding"), ("edge_emb", "edge_embedding"),
                            ("edge_head", "edge_classification"), ("node_head", "node_classification"),
                            ("class_head", "classification")):
            setattr(p, field, _mlp_train_struct(getattr(model, name), name, offsets))
        p.w1, p.b1 = offsets["mpn_node_cls.mlp_edge.0.weight"], offsets["mpn_node_cls.mlp_edge.0.bias"]
        p.w2, p.b2 = offsets["mpn_node_cls.mlp_edge.2.weight"], offsets["mpn_node_cls.mlp_edge.2.bias"]
        nt = None
        if model.aggr_type == "per_type":          # TypeAwareMPNLayer: 17 message matrices, attention, update over all types
            nt = node_types.detach().contiguous()
            p.per_type, p.num_types, p.attn = 1, layer.num_types, nv.ATTN[layer.aggr_sub]
            p.node_types = nt.data_ptr()
            p.wm, p.bm = offsets["mpn_node_cls.mlp_node.mlp.0.0.weight"], offsets["mpn_node_cls.mlp_node.mlp.0.0.bias"]
            p.wm_type_stride = offsets["mpn_node_cls.mlp_node.mlp.1.0.weight"] - p.wm
            if layer.attn_net is not None:
                p.wa, p.ba = offsets["mpn_node_cls.attn_net.0.weight"], offsets["mpn_node_cls.attn_net.0.bias"]
        else:
            p.wm, p.bm = offsets["mpn_node_cls.mlp_node.0.weight"], offsets["mpn_node_cls.mlp_node.0.bias"]
        if layer.update_mlp is not None:
            p.wu, p.bu = offsets["mpn_node_cls.update_mlp.0.weight"], offsets["mpn_node_cls.update_mlp.0.bias"]
        with torch.cuda.device(dev):
            ws_bytes = int(lib.pgmp_mpn_train_workspace_bytes(p))
            # The activations of all steps live in this workspace until the reverse pass (2.6 GB at 8 images): it is
            # handed back to a per-module pool after backward() instead of going through the allocator every step.
            pool = model.__dict__.setdefault("_train_ws_pool", [])
            stream = torch.cuda.current_stream().cuda_stream       # a pooled workspace is only reused on the stream it was last used on
            hit = next((e for e in pool if e[0] == stream and e[1].device == dev and e[1].numel() >= ws_bytes), None)
            if hit is not None:
                pool.remove(hit)
                ws = hit[1]
            else:
                pool.clear()
                ws = torch.empty(ws_bytes + ws_bytes // 8, dtype=torch.uint8, device=dev)
            p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
            ctx.token = object()                                    # this forward owns the workspace until another forward takes it
            owners = model.__dict__.setdefault("_train_ws_owner", {})
            owners[ws.data_ptr()] = ctx.token
            # the reference evaluates the node and class heads once more on the final features (:93-94): with BatchNorm in
            # the heads (MPN.BN) that second call applies the momentum update of their running statistics a second time
            twice = [m for head in (model.node_classification, model.classification) for m in head
                     if isinstance(m, nn.BatchNorm1d)]
            before = [(m.running_mean.clone(), m.running_var.clone()) for m in twice]
            nv.check(lib.pgmp_mpn_train_forward(p, nv.current_stream()))
            for m, (mean0, var0) in zip(twice, before):      # r2 = (1 - mom) r1 + mom b  with  mom b = r1 - (1 - mom) r0
                keep = 1.0 - m.momentum
                m.running_mean.mul_(1.0 + keep).sub_(mean0, alpha=keep)
                m.running_var.mul_(1.0 + keep).sub_(var0, alpha=keep)
                m.num_batches_tracked += 1
        for mod in model.modules():            # the kernels updated the running statistics in place
            if isinstance(mod, nn.BatchNorm1d):
                torch.autograd.graph.increment_version((mod.running_mean, mod.running_var))
                mod.num_batches_tracked += 1
        ctx.p, ctx.sizes, ctx.shapes = p, sizes, [tuple(q.shape) for q in params]
        ctx.offsets = [offsets[n] for n in names]
        ctx.keep = (flat, x_, ea, ei, ws, edge_logits, node_logits, class_logits) + ((nt,) if nt is not None else ())   # the forward's activations live in ws
        ctx.pool, ctx.owners = pool, owners
        ctx.need_x = x.requires_grad
        return edge_logits, node_logits, class_logits

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_edge, d_node, d_class):
        p = ctx.p
        flat, x_ = ctx.keep[0], ctx.keep[1]
        if ctx.owners.get(ctx.keep[4].data_ptr()) is not ctx.token:
            raise RuntimeError("the activations of this forward pass are gone: its workspace went back to the module's pool "
                               "after the first backward() and a later forward() has reused it")
        dev = flat.device
        grads = torch.zeros_like(flat)
        grad_x = torch.empty_like(x_) if ctx.need_x else None
        de, dn, dc = d_edge.contiguous().float(), d_node.contiguous().float(), d_class.contiguous().float()
        p.grads, p.grad_x = grads.data_ptr(), (grad_x.data_ptr() if grad_x is not None else None)
        p.d_edge_logits, p.d_node_logits, p.d_class_logits = de.data_ptr(), dn.data_ptr(), dc.data_ptr()
        with torch.cuda.device(dev):
            nv.check(nv.lib().pgmp_mpn_train_backward(p, nv.current_stream()))
            for t in ctx.keep + (de, dn, dc):
                t.record_stream(torch.cuda.current_stream())
        out = [grads[off:off + k].view(shape) for off, k, shape in zip(ctx.offsets, ctx.sizes, ctx.shapes)]
        if len(ctx.pool) < 2 and not any(e[1] is ctx.keep[4] for e in ctx.pool):
            ctx.pool.append((torch.cuda.current_stream().cuda_stream, ctx.keep[4]))   # stream-ordered reuse
        return (None, grad_x, None, None, None) + tuple(out)


class NodeClassificationMPNSimple(nn.Module):
    """Same constructor, parameters (names and shapes) and ``forward`` contract as the reference class
    (NodeClassificationMPNSimple.py:23-97); inference (``eval()``) runs in ``libpgmp.so``."""

    def __init__(self, config):
        super().__init__()
        self.use_skip_connections = config.SKIP
        self.node_summary = config.NODE_TYPE_SUMMARY
        if not (config.NODE_FEATURE_DIM == config.EDGE_FEATURE_DIM == config.EDGE_FEATURE_HIDDEN == 64):
            raise NotImplementedError("NODE_FEATURE_DIM / EDGE_FEATURE_DIM / EDGE_FEATURE_HIDDEN must be 64 "
                                      "(every NodeClassificationMPN config of the reference)")
        if config.LATE_FUSION_POS:
            raise NotImplementedError("LATE_FUSION_POS is out of scope")
        self.aggr_type = config.AGGR_TYPE
        if config.AGGR_TYPE == "agnostic":
            self.mpn_node_cls = MPLayer(config.NODE_FEATURE_DIM, config.EDGE_FEATURE_DIM, config.EDGE_FEATURE_HIDDEN,
                                        aggr=config.AGGR, skip=config.SKIP,
                                        use_node_update_mlp=config.USE_NODE_UPDATE_MLP, edge_mlp=config.EDGE_MLP)
        elif config.AGGR_TYPE == "per_type":
            num_types = {"per_body_part": 6, "not": config.NUM_JOINTS, "left_right": 9}.get(self.node_summary)
            self.mpn_node_cls = TypeAwareMPNLayer(config.NODE_FEATURE_DIM, config.EDGE_FEATURE_DIM,
                                                  config.EDGE_FEATURE_HIDDEN, aggr=config.AGGR, skip=config.SKIP,
                                                  edge_mlp=config.EDGE_MLP, num_types=num_types,
                                                  aggr_sub=config.AGGR_SUB, update_type=config.UPDATE_TYPE)
        else:
            raise NotImplementedError("AGGR_TYPE=%r" % (config.AGGR_TYPE,))
        if config.AGGR not in nv.AGGR:
            raise NotImplementedError("AGGR=%r" % (config.AGGR,))
        self.edge_embedding = make_mlp(config.EDGE_INPUT_DIM, config.EDGE_EMB.OUTPUT_SIZES, bn=config.EDGE_EMB.BN,
                                       end_with_relu=config.EDGE_EMB.END_WITH_RELU)
        self.node_embedding = make_mlp(config.NODE_INPUT_DIM, config.NODE_EMB.OUTPUT_SIZES, bn=config.NODE_EMB.BN,
                                       end_with_relu=config.NODE_EMB.END_WITH_RELU)
        self.edge_classification = make_mlp(config.EDGE_FEATURE_DIM, config.EDGE_CLASS.OUTPUT_SIZES, bn=config.BN)
        self.node_classification = make_mlp(config.NODE_FEATURE_DIM, config.NODE_CLASS.OUTPUT_SIZES, bn=config.BN)
        self.classification = make_mlp(config.NODE_FEATURE_DIM, config.CLASS.OUTPUT_SIZES, bn=config.BN)
        self.edge_steps = config.STEPS
        self.node_steps = config.NODE_STEPS
        self.aux_loss_steps = config.AUX_LOSS_STEPS
        self.aggr = config.AGGR
        self.precision = getattr(config, "B200_PRECISION", "fp32")
        if self.precision not in nv.PRECISION:
            raise ValueError("B200_PRECISION must be one of %s" % (sorted(nv.PRECISION),))
        self._pack_cache = None
        self._status = None          # (pinned status word of the previous forward, event)

    # ------------------------------------------------------------------ weight packing
    def invalidate_packed_weights(self):
        """Drop the packed (BatchNorm-folded, bf16-split) copy of the weights.  The cache notices ``load_state_dict``,
        optimizer steps and every other in-place update that goes through autograd's version counter, and re-allocated
        storage; updates through ``.data`` (``p.data.copy_()``, EMA swaps) bump no counter -- call this after them."""
        self._pack_cache = None

    def check_status(self):
        """Raise if the previous ``forward`` met malformed input (an ``edge_index`` entry outside ``[0, N)``: the reference
        raises an IndexError).  The forward itself never waits for the device; the status word is read here and at the
        start of the next forward, once its copy has arrived."""
        st = self._status
        if st is not None and st[1].query():
            self._status = None
            if int(st[0][0]) & 1:
                raise IndexError("edge_index of the previous forward held node ids outside [0, num_nodes): those edges were "
                                 "dropped and their logits are undefined")

    def _pack(self, device):
        # the (container, name) slots of every parameter / buffer are collected once: walking the module tree through
        # parameters() + buffers() on every forward cost 0.6 ms of host time per batch (a sixth of the GPU time of a
        # 32-image batch); a re-assigned Parameter is still seen, the slots are looked up each time
        slots = self.__dict__.get("_pack_slots")
        if slots is None:
            slots = [(d, n) for m in self.modules() for d in (m._parameters, m._buffers) for n in d]
            self.__dict__["_pack_slots"] = slots
        key = (str(device),) + tuple((t._version, t.data_ptr()) for t in (d[n] for d, n in slots) if t is not None)
        if self._pack_cache is not None and self._pack_cache[0] == key:
            return self._pack_cache[1]
        pk = _Packer()
        layer = self.mpn_node_cls
        skip = bool(self.use_skip_connections)
        nd = 128 if skip else 64
        spec = {name: _fold_mlp(getattr(self, name), pk) for name in
                ("node_embedding", "edge_embedding", "edge_classification", "node_classification", "classification")}
        W1 = layer.mlp_edge[0].weight.detach().float()            # [64, 2*nd + ed], input = [x_i ; x_j ; e] (layers.py:214)
        off = {"w1_dst": pk.add(W1[:, :nd].t()), "w1_src": pk.add(W1[:, nd:2 * nd].t())}
        if skip:                                                   # e = [e_initial ; e_current] (NodeClassificationMPNSimple.py:78)
            off["w1_e0"] = pk.add(W1[:, 2 * nd:2 * nd + 64].t())
            off["w1_e"] = pk.add(W1[:, 2 * nd + 64:].t())
        else:
            off["w1_e"] = pk.add(W1[:, 2 * nd:].t())
        off["b1"] = pk.add(layer.mlp_edge[0].bias)
        off["w2"] = pk.add(layer.mlp_edge[2].weight.detach().float().t())
        off["b2"] = pk.add(layer.mlp_edge[2].bias)
        per_type = isinstance(layer, TypeAwareMPNLayer)
        lins = [m[0] for m in layer.mlp_node.mlp] if per_type else [layer.mlp_node[0]]   # input = [x_i ; e'] (layers.py:223,273)
        off["wm_x"] = pk.add(torch.stack([l.weight.detach().float()[:, :nd].t() for l in lins]))
        off["wm_e"] = pk.add(torch.stack([l.weight.detach().float()[:, nd:].t() for l in lins]))
        off["bm"] = pk.add(torch.stack([l.bias.detach().float() for l in lins]))
        attn_net = getattr(layer, "attn_net", None)
        if attn_net is not None:
            off["wa"] = pk.add(attn_net[0].weight.detach().float().t())
            off["ba"] = pk.add(attn_net[0].bias)
        hier = per_type and getattr(layer, "update_type", "mlp") == "hierarch_mlp"
        if hier:      # layers.py:89-128: [out][in] weights and biases back to back: 7 first, 6 second, final
            um = layer.update_mlp
            off["hier"] = pk.add(torch.cat([t.detach().float().reshape(-1) for lin in
                                            (list(um.first_layer) + list(um.second_layer) + [um.final])
                                            for t in (lin.weight, lin.bias)]))
        elif layer.update_mlp is not None:
            off["wu"] = pk.add(layer.update_mlp[0].weight.detach().float().t())
            off["bu"] = pk.add(layer.update_mlp[0].bias)
        flat = pk.finish(device)

        def split(w):      # bf16x3 operands of the tensor-core path: hi = bf16(W), lo = bf16(W - hi), [out][in]
            w = w.detach().float().to(device)
            hi = w.to(torch.bfloat16)
            lo = (w - hi.float()).to(torch.bfloat16)
            return torch.stack([hi, lo])

        def tile_images(w):    # [2][64][K] hi/lo -> [K/64][2][64][64] SWIZZLE_128B tile images (chunk c of row r at c ^ (r & 7))
            k_blocks = w.shape[2] // 64
            t = w.reshape(2, 64, k_blocks, 8, 8).permute(2, 0, 1, 3, 4)          # [kb][2][row][chunk][8]
            r = torch.arange(64, device=w.device)[:, None]
            src = (torch.arange(8, device=w.device)[None, :] ^ (r & 7))            # position p holds chunk p ^ (r & 7)
            idx = src[None, None, :, :, None].expand(k_blocks, 2, 64, 8, 8)
            return torch.gather(t, 3, idx).reshape(k_blocks, 2, 64, 64)

        w1e = W1[:, 2 * nd + 64:] if skip else W1[:, 2 * nd:]
        tc = dict(tc_w1_e=split(w1e).contiguous(), tc_w2=split(layer.mlp_edge[2].weight).contiguous(),
                  tc_wm_e=torch.stack([split(l.weight[:, nd:]) for l in lins]).contiguous(),
                  tc_wtab=torch.stack([tile_images(split(W1[:, :nd])), tile_images(split(W1[:, nd:2 * nd]))] +
                                      [tile_images(split(l.weight[:, :nd])) for l in lins]).contiguous())
        nemb_w = spec["node_embedding"][2]          # folded Linear weights of the node embedding
        if [tuple(w_.shape) for w_ in nemb_w] == [(128, 128), (64, 128), (64, 64)]:
            tc["tc_wnemb"] = torch.cat([split(w_).reshape(-1) for w_ in nemb_w]).contiguous()
        emb_w = spec["edge_embedding"][2]           # folded Linear weights of the edge embedding
        if all(max(w.shape) <= 64 for w in emb_w):
            pad = torch.zeros(len(emb_w), 2, 64, 64, dtype=torch.bfloat16, device=device)
            for l, w_ in enumerate(emb_w):
                pad[l, :, :w_.shape[0], :w_.shape[1]] = split(w_)
            tc["tc_wemb"] = pad.contiguous()
        if skip:
            tc["tc_w1_e0"] = split(W1[:, 2 * nd:2 * nd + 64]).contiguous()
        head_w = spec["edge_classification"][2]     # folded Linear weights of the edge head
        if [tuple(w.shape) for w in head_w] == [(64, 64), (32, 64), (1, 32)]:
            tc["tc_wh1"], tc["tc_wh2"] = split(head_w[0]).contiguous(), split(head_w[1]).contiguous()
        nh_w, ch_w = spec["node_classification"][2], spec["classification"][2]
        if ([tuple(w_.shape) for w_ in nh_w[:2]] == [(64, 64), (32, 64)]
                and [tuple(w_.shape) for w_ in ch_w[:2]] == [(64, 64), (32, 64)]):
            tc["tc_wheads"] = torch.cat([split(w_).reshape(-1) for w_ in (nh_w[0], ch_w[0], nh_w[1], ch_w[1])]).contiguous()
        if layer.update_mlp is not None and not hier:
            wu = layer.update_mlp[0].weight.detach().float()
            tc["tc_wu"] = torch.stack([split(wu[:, t * 64:(t + 1) * 64]) for t in range(wu.shape[1] // 64)]).contiguous()
        packed = dict(flat=flat, tc=tc, spec=spec, off=off, per_type=per_type, skip=skip, hier=hier,
                      num_types=layer.num_types if per_type else 1,
                      attn=nv.ATTN[layer.aggr_sub] if per_type else 0,
                      num_classes=self.classification[-1].out_features)
        self._pack_cache = (key, packed)
        return packed

    def num_outputs(self):
        first = max(self.edge_steps - self.aux_loss_steps - 1, 0)       # NodeClassificationMPNSimple.py:81
        return self.edge_steps - first

    # ------------------------------------------------------------------ forward
    def forward(self, x, edge_attr, edge_index, **kwargs):
        if self.training:
            return self._forward_train(x, edge_attr, edge_index, **kwargs)
        self.check_status()
        if self.node_steps != 0:
            raise NotImplementedError("NODE_STEPS != 0 (the reference's own loop omits node_types, App. A)")
        if self.edge_steps < 1:
            raise NotImplementedError("STEPS must be >= 1")
        nv.require_cuda(x, "x", torch.float32)
        nv.require_cuda(edge_attr, "edge_attr", torch.float32)
        nv.require_cuda(edge_index, "edge_index", torch.int64)
        node_types = sum_node_types(self.node_summary, kwargs["node_types"])          # :64
        nv.require_cuda(node_types, "node_types", torch.int64)
        dev = x.device
        N, E = x.shape[0], edge_index.shape[1]
        n_out = self.num_outputs()
        pk = self._pack(dev)
        J = pk["num_classes"]
        if edge_attr.shape[0] != E or (E and edge_attr.shape[1] != self.edge_embedding[0].in_features):
            raise ValueError("edge_attr must be [E, %d]" % self.edge_embedding[0].in_features)
        if x.shape[1] != self.node_embedding[0].in_features or node_types.shape[0] != N:
            raise ValueError("x must be [N, %d] and node_types [N]" % self.node_embedding[0].in_features)
        edge_logits = torch.empty((n_out, E), dtype=torch.float32, device=dev)
        node_logits = torch.empty((n_out, N), dtype=torch.float32, device=dev)
        class_logits = torch.empty((n_out, N, J), dtype=torch.float32, device=dev)
        if N > 0:
            lib = nv.lib()
            x_ = x.detach()
            ea = edge_attr.detach().contiguous()
            ei = edge_index.detach().contiguous()
            nt = node_types.detach().contiguous()
            base = pk["flat"].data_ptr()
            p = nv.MpnParams(num_nodes=N, num_edges=E, x=x_.data_ptr(), x_stride_n=x_.stride(0),
                             x_stride_c=x_.stride(1), edge_attr=ea.data_ptr(), edge_index=ei.data_ptr(),
                             node_types=nt.data_ptr(), dim=64, per_type=int(pk["per_type"]),
                             num_types=pk["num_types"], num_type_mlps=17 if pk["per_type"] else 1,
                             skip=int(pk["skip"]), steps=self.edge_steps, aux_loss_steps=self.aux_loss_steps,
                             aggr=nv.AGGR[self.aggr], attn=pk["attn"],
                             has_update_mlp=int(self.mpn_node_cls.update_mlp is not None), update_hier=int(pk["hier"]),
                             num_classes=J,
                             precision=nv.PRECISION[self.precision])
            for field, name in (("node_emb", "node_embedding"), ("edge_emb", "edge_embedding"),
                                ("edge_head", "edge_classification"), ("node_head", "node_classification"),
                                ("class_head", "classification")):
                setattr(p, field, _mlp_struct(pk["spec"][name], base))
            for k, o in pk["off"].items():
                setattr(p, k, base + 4 * o)
            for k, t in pk["tc"].items():
                setattr(p, k, t.data_ptr())
            p.edge_logits, p.node_logits, p.class_logits = edge_logits.data_ptr(), node_logits.data_ptr(), class_logits.data_ptr()
            with torch.cuda.device(dev):
                ws_bytes = int(lib.pgmp_mpn_workspace_bytes(p))
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                p.workspace, p.workspace_bytes = ws.data_ptr(), ws_bytes
                nv.check(lib.pgmp_mpn_forward(p, nv.current_stream()))
                status_h = torch.empty(1, dtype=torch.int32, pin_memory=True)
                status_h.copy_(ws[:4].view(torch.int32), non_blocking=True)     # first word of the workspace (pgmp.h)
                ev = torch.cuda.Event()
                ev.record()
                self._status = (status_h, ev)
                ws.record_stream(torch.cuda.current_stream())
        # lists of independent tensors: the caller mutates entries in place (PoseEstimation.py:95-101);
        # .squeeze() semantics of :82,:84,:93 (0-d when there is a single edge / node)
        preds_edge = [edge_logits[i].squeeze() for i in range(n_out)]
        preds_node = [node_logits[i].squeeze() for i in range(n_out)]
        preds_class = [class_logits[i] for i in range(n_out)]
        preds_node.append(preds_node[-1].clone())                      # :93-94 (recomputed on the same features)
        preds_class.append(preds_class[-1].clone())
        return preds_edge, preds_node, preds_class, [None]


def _forward_train(self, x, edge_attr, edge_index, **kwargs):
    """``train()`` mode: BatchNorm batch statistics, activations kept for the reverse pass."""
    node_types = None
    if self.aggr_type == "per_type":
        if self.mpn_node_cls.update_type != "mlp":
            raise NotImplementedError("training kernels cover UPDATE_TYPE mlp; %r trains with the reference for now"
                                      % (self.mpn_node_cls.update_type,))
        node_types = sum_node_types(self.node_summary, kwargs["node_types"])          # :64
        nv.require_cuda(node_types, "node_types", torch.int64)
        if node_types.shape[0] != x.shape[0]:
            raise ValueError("node_types must be [N]")
    if self.node_steps != 0 or self.edge_steps < 1:
        raise NotImplementedError("NODE_STEPS != 0 / STEPS < 1")
    nv.require_cuda(x, "x", torch.float32)
    nv.require_cuda(edge_attr, "edge_attr", torch.float32)
    nv.require_cuda(edge_index, "edge_index", torch.int64)
    N, E = x.shape[0], edge_index.shape[1]
    if N == 0:
        raise ValueError("empty graph")
    if edge_attr.shape[0] != E or (E and edge_attr.shape[1] != self.edge_embedding[0].in_features):
        raise ValueError("edge_attr must be [E, %d]" % self.edge_embedding[0].in_features)
    if x.shape[1] != self.node_embedding[0].in_features:
        raise ValueError("x must be [N, %d]" % self.node_embedding[0].in_features)
    params = [p for _, p in self.named_parameters()]
    for p in params:
        nv.require_cuda(p, "parameter", torch.float32)
    edge_logits, node_logits, class_logits = _MpnTrainFunction.apply(self, x, edge_attr, edge_index, node_types, *params)
    n_out = self.num_outputs()
    preds_edge = [edge_logits[i].squeeze() for i in range(n_out)]
    preds_node = [node_logits[i].squeeze() for i in range(n_out)]
    preds_class = [class_logits[i] for i in range(n_out)]
    preds_node.append(preds_node[-1].clone())                          # :93-94
    preds_class.append(preds_class[-1].clone())
    return preds_edge, preds_node, preds_class, [None]


NodeClassificationMPNSimple._forward_train = _forward_train


_OUT_OF_SCOPE = (
    "VanillaMPN", "ClassificationMPN", "ClassificationMPNSimple", "VanillaMPN2", "ClassificationNaive",
    "NodeClassificationMPNWithBackground", "NodeClassificationMPNTypeBased", "NodeClassificationMPNAttention",
    "NodeClassificationMPNSelfAttention", "NodeClassificationMPNWithRef", "NodeClassificationMPNFPConstrained",
    "NodeClassificationMPNTypeConstrained", "NodeClassificationMPNTag", "MPNTag", "LogisticEdgeClassifier",
    "NodeClassificationMPNGroupBased", "NodeClassificationMPNGroupBasedHierach", "JointTypeClassification",
    "TagThreshold", "PlainTag")


def get_mpn_model(config, **kwargs):
    """src/Models/MessagePassingNetwork/__init__.py:27-73."""
    if config.NAME == "NodeClassificationMPN":
        return NodeClassificationMPNSimple(config)
    if config.NAME in _OUT_OF_SCOPE:
        raise NotImplementedError("MPN variant %r is outside the B200 hot path (SURVEY.md 2, row 2); use the "
                                  "reference implementation for it" % (config.NAME,))
    raise NotImplementedError(config.NAME)
