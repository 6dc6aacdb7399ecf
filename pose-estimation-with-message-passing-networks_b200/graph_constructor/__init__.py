"""Drop-in for the reference's ``graph_constructor`` package
(``src/graph_constructor/__init__.py:4-5``, ``ConstructGraph.py:9-249``): inference, and training with
``joints_gt`` / ``factor_list`` (label slots of the 15-tuple, ``labels.py``).

``get_graph_constructor(config, **kwargs).construct_graph()`` keeps the reference's
constructor keywords and its 15-tuple; the work runs in ``libpgmp.so``
(``csrc/gc.cu``) for the whole batch at once with a single host read (the counts
that size the outputs) instead of the reference's per-image Python loop.
"""

import torch

from .. import _native as nv

KNN_K = 50            # ConstructGraph.py:365
NO_THRESHOLD_K = 20   # ConstructGraph.py:1185


class _GatherNodeFeatures(torch.autograd.Function):
    """x = features[b, :, y, x] with gradients flowing back into ``features``
    (end-to-end training, train.py:232); forward is the CUDA gather inside ``pgmp_gc_emit``."""

    @staticmethod
    def forward(ctx, features, x_out, batch_index, joint_det):
        ctx.save_for_backward(batch_index, joint_det)
        ctx.feat_shape = features.shape
        return x_out.view_as(x_out)

    @staticmethod
    def backward(ctx, grad):
        batch_index, joint_det = ctx.saved_tensors
        g = torch.zeros(ctx.feat_shape, dtype=torch.float32, device=grad.device)     # the dense gradient the backbone expects
        gx = grad.contiguous().float()
        with torch.cuda.device(grad.device):
            nv.check(nv.lib().pgmp_gc_gather_backward(gx.data_ptr(), joint_det.data_ptr(), batch_index.data_ptr(), gx.shape[0],
                                                      gx.shape[1], g.data_ptr(), g.stride(0), g.stride(1), g.stride(2),
                                                      g.stride(3), nv.current_stream()))
        return g, None, None, None


class _GatherConvFeatures(torch.autograd.Function):
    """``ConvUpsampleFeatures`` under autograd (end-to-end training, train.py:232): the forward is the fused CUDA kernel
    inside ``construct_graph``; the reverse pass returns the gradients of the backbone map, of the ``feature_gather``
    weight and of its bias.  ``x = bias + P W`` with the interpolated input patches ``P [N, 9 Cin]``: ``d W = P^T d x`` and
    ``d P = d x W^T`` are plain fp32 library products, the patches and the transposed interpolation back into the map are
    ``pgmp_gc_gather_conv_patches`` / ``pgmp_gc_gather_conv_backward`` (no atomics, fixed summation order)."""

    @staticmethod
    def forward(ctx, feat, weight, bias, x_out, batch_index, joint_det, size):
        ctx.save_for_backward(feat, weight, batch_index, joint_det)
        ctx.size = size
        return x_out.view_as(x_out)

    @staticmethod
    def backward(ctx, grad):
        feat, weight, batch_index, joint_det = ctx.saved_tensors
        lib = nv.lib()
        gx = grad.contiguous().float()
        N, (cout, cin) = gx.shape[0], weight.shape[:2]
        fm = feat.detach()

        def params(t):
            return nv.GatherConvParams(
                features=t.data_ptr(), feat_stride_b=t.stride(0), feat_stride_c=t.stride(1), feat_stride_y=t.stride(2),
                feat_stride_x=t.stride(3), cin=cin, height=fm.shape[2], width=fm.shape[3], cout=cout, out_height=ctx.size[0],
                out_width=ctx.size[1], joint_det=joint_det.data_ptr(), batch_index=batch_index.data_ptr(), num_nodes=N)

        d_feat = d_weight = d_bias = None
        with torch.cuda.device(gx.device):
            if ctx.needs_input_grad[1]:
                patches = torch.empty((N, 9 * cin), dtype=torch.float32, device=gx.device)
                nv.check(lib.pgmp_gc_gather_conv_patches(params(fm), patches.data_ptr(), nv.current_stream()))
                d_weight = (patches.t() @ gx).view(3, 3, cin, cout).permute(3, 2, 0, 1).contiguous()
            if ctx.needs_input_grad[2]:
                d_bias = gx.sum(0)
            if ctx.needs_input_grad[0]:
                wt = weight.detach().float().permute(2, 3, 1, 0).reshape(9 * cin, cout)      # [(ky, kx, ci), co]
                d_patches = (gx @ wt.t()).contiguous()
                d_feat = torch.zeros(fm.shape, dtype=torch.float32, device=gx.device)
                nv.check(lib.pgmp_gc_gather_conv_backward(params(d_feat), d_patches.data_ptr(), fm.shape[0], d_feat.data_ptr(),
                                                          nv.current_stream()))
        return d_feat, d_weight, d_bias, None, None, None, None


def _host_resident(t, need_contiguous=False):
    """A pinned fp32 host tensor the gather kernels can read in place (no gradient flows into it)."""
    return (t is not None and t.device.type == "cpu" and t.is_pinned() and t.dtype == torch.float32
            and not t.requires_grad and (t.is_contiguous() or not need_contiguous))


class ConvUpsampleFeatures:
    """Lazy ``interpolate(feature_gather(feat), size, mode="bilinear", align_corners=False)``.

    Pass it as ``features=`` instead of the materialised ``[B, 128, H, W]`` tensor: the graph constructor then
    evaluates the 3x3 convolution and the interpolation only at the candidate pixels, straight from the backbone's
    feature map (``PoseEstimation.py:64-66, 79, 341, 442-450`` + ``ConstructGraph.py:265, 269``).

    ``feat``: ``[B, Cin, h, w]`` float32, CUDA or pinned host.  ``conv``: the model's ``feature_gather``
    (``nn.Conv2d(Cin, Cout <= 128, 3, 1, 1)``) or a ``(weight [Cout,Cin,3,3], bias [Cout])`` pair.  ``size``: the
    ``(H, W)`` the reference interpolates to (the heatmap size; equal to ``(h, w)`` in the training ``forward``).
    Under autograd (``feat`` or the convolution's parameters require a gradient) the reverse pass is native too
    (``_GatherConvFeatures``): the ``[B, Cout, H, W]`` maps are not materialised in training either, and the
    ``feature_gather`` gradients join the MPN's in ``parallel.allreduce_gradients``.
    """

    def __init__(self, feat, conv, size):
        weight, bias = (conv.weight, conv.bias) if hasattr(conv, "weight") else conv
        if feat.dim() != 4 or feat.dtype != torch.float32:
            raise TypeError("feat must be a float32 [B, Cin, h, w] tensor")
        if tuple(weight.shape[1:]) != (feat.shape[1], 3, 3):
            raise ValueError("feature_gather must be a 3x3 convolution over %d channels (PoseEstimation.py:64)" % feat.shape[1])
        if hasattr(conv, "padding") and (tuple(conv.padding) != (1, 1) or tuple(conv.stride) != (1, 1)):
            raise NotImplementedError("FEATURE_GATHER_PADDING / stride other than 1 (default_config.py:26-27)")
        if weight.shape[0] > 128:
            raise NotImplementedError("more than 128 output channels")
        self.feat = feat
        self.weight, self.bias = weight, bias
        self.size = (int(size[0]), int(size[1]))
        self.shape = (feat.shape[0], weight.shape[0]) + self.size
        self.dtype = torch.float32
        self.requires_grad = False

    def pack(self, device):
        """Transposed weights ``[(ky, kx, ci), co]`` and bias on ``device`` (cached per parameter version)."""
        key = (self.weight.data_ptr(), self.weight._version, str(device))
        if getattr(self, "_packed_key", None) != key:
            w = self.weight.detach().to(device=device, dtype=torch.float32)
            self._wt = w.permute(2, 3, 1, 0).reshape(-1, w.shape[0]).contiguous()
            b = self.bias if self.bias is not None else torch.zeros(w.shape[0])
            self._b = b.detach().to(device=device, dtype=torch.float32).contiguous()
            self._packed_key = key
        return self._wt, self._b


class HeadStages:
    """Lazy stand-in for the ``scoremaps`` argument: the two output stages of the HigherHRNet head, assembled INSIDE the
    NMS kernel's load stage (``pgmp_gc_detect_fused``) instead of by ``hr_process_output`` in front of it
    (src/Models/HigherHRNet/hrnet.py:587-611: interpolate the half-resolution stage, average with the full-resolution
    one).  The assembled ``[B, J, H, W]`` map -- 570 MB at 32 x 17 x 512 x 512 -- is then neither written nor re-read;
    ``keep_scoremaps=True`` still writes it (``.scoremaps`` after the detection) for the pose-assembly tail
    (``refine`` / ``adjust``).  ``flipped`` / ``flip_index`` add the flip-test average of one scale
    (PoseEstimation.py:343-402, multi_scales_testing.py:162): ``(A + A_flipped[:, flip_index, :, ::-1]) / 2``.
    Results are bit-identical to ``hr_process_output`` + the plain constructor.  The same object can be passed as
    ``tagmaps``: the detections' tags are then interpolated from the half-resolution stage at their pixels only
    (``pgmp_gc_gather_stage_tags``) and the up-sampled tag maps are not built either."""

    def __init__(self, outputs, num_joints, mode="avg", flipped=None, flip_index=None, keep_scoremaps=False):
        s1, s2 = outputs
        if mode not in ("avg", "small"):
            raise NotImplementedError("mode %r ('large' needs no assembly: pass scoremap_2)" % (mode,))
        if flipped is not None and flip_index is None:
            raise ValueError("flipped outputs need flip_index (FLIP_CONFIG)")
        self.terms = [(s1, s2)] + ([tuple(flipped)] if flipped is not None else [])
        for a, b in self.terms:
            nv.require_cuda(a, "scoremap_1", torch.float32)
            nv.require_cuda(b, "scoremap_2", torch.float32)
            if a.shape != s1.shape or b.shape != s2.shape:
                raise ValueError("flipped outputs must have the shapes of the plain ones")
        B, C1 = s1.shape[0], s1.shape[1]
        if C1 < num_joints or s2.shape[0] != B or (mode == "avg" and s2.shape[1] != num_joints):
            raise ValueError("scoremap_1 must be [B, >= J, h, w] and scoremap_2 [B, J, H, W]")
        if num_joints > 32:
            raise NotImplementedError("more than 32 joint types")
        self.mode, self.num_joints = mode, num_joints
        self.flip_index = [int(i) for i in flip_index] if flip_index is not None else None
        if self.flip_index is not None and sorted(self.flip_index) != list(range(num_joints)):
            raise ValueError("flip_index must be a permutation of the joint types")
        self.keep_scoremaps = keep_scoremaps
        self.scoremaps = None
        self.shape = (B, num_joints, s2.shape[2], s2.shape[3])
        self.device = s1.device
        self.dtype = torch.float32

    def dim(self):
        return 4

    def assembly(self, out):
        """The C-ABI description of the terms (tensors made contiguous are kept alive on ``self``)."""
        self._keep = [(a.detach().contiguous(), b.detach().contiguous()) for a, b in self.terms]
        a = nv.GcAssembly()
        for t, (x1, x2) in enumerate(self._keep):
            a.stage1[t], a.stage2[t] = x1.data_ptr(), x2.data_ptr()
        a.channels1, a.h, a.w = self._keep[0][0].shape[1:4]
        a.mode, a.n_terms = (0 if self.mode == "avg" else 1), len(self._keep)
        for j, i in enumerate(self.flip_index or []):
            a.flip_index[j] = i
        a.scoremaps_out = out.data_ptr() if out is not None else None
        return a


class NaiveGraphConstructor:
    """Same constructor signature as the reference class (ConstructGraph.py:11)."""

    def __init__(self, scoremaps, tagmaps, features, joints_gt, factor_list, masks, device, config, testing,
                 heatmaps, num_joints):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("pgmp_b200 graph constructor needs a CUDA device (no CPU fallback)")
        # pinned host heatmaps are copied by the detection launch itself (on its stream, non-blocking): with
        # detect_async() the copy of the next batch overlaps the current batch's kernels
        if isinstance(scoremaps, HeadStages):
            self.scoremaps = scoremaps                       # assembled inside the NMS kernel (pgmp_gc_detect_fused)
        else:
            self.scoremaps = scoremaps if _host_resident(scoremaps, need_contiguous=True) else scoremaps.to(self.device)
        self._pending = None
        # The reference moves every input to the device (ConstructGraph.py:12-18).  Only N pixels of the feature and
        # tag maps are ever read (N x C x 4 bytes of a 134 MB map per image), so pinned host tensors are left where
        # they are and the gather kernels read those pixels in place over PCIe (unified addressing); like a
        # non_blocking copy from pinned memory, the caller must not overwrite them before the stream has caught up.
        if isinstance(tagmaps, HeadStages):
            self.tagmaps = tagmaps                           # tags evaluated at the detections (pgmp_gc_gather_stage_tags)
        else:
            self.tagmaps = tagmaps if _host_resident(tagmaps, need_contiguous=True) else (
                tagmaps.to(self.device) if tagmaps is not None else None)
        if isinstance(features, ConvUpsampleFeatures):
            # every candidate reads a 4 x 4 x Cin neighbourhood of the (small) backbone map: a host-resident map is
            # copied, reading it in place would cost more PCIe transactions than the copy
            if features.feat.device.type != "cuda":
                features.feat = features.feat.to(self.device, non_blocking=True)
            self.features = features
        else:
            self.features = features if _host_resident(features) else (
                features.to(self.device) if features is not None else None)
        self.joints_gt = joints_gt
        self.factor_list = factor_list
        self.masks = masks.to(self.device) if masks is not None else None
        self.batch_size = scoremaps.shape[0]
        self.num_joints = num_joints
        self.testing = testing
        self.config = config

        if config.USE_GT or config.CHEAT:
            raise NotImplementedError("USE_GT / CHEAT (graphs built from the ground truth, ConstructGraph.py:70-99) are "
                                      "debugging modes outside the hot path")
        # training: labels from the ground truth (ConstructGraph.py:104-168), graph_constructor/labels.py
        self.edge_label_method = config.EDGE_LABEL_METHOD
        self.include_neighbouring_keypoints = config.USE_NEIGHBOURS
        self.matching_radius = config.MATCHING_RADIUS
        self.inclusion_radius = config.INCLUSION_RADIUS
        self.with_background_class = config.WITH_BACKGROUND
        self.node_dropout = config.NODE_DROPOUT if config.NODE_DROPOUT != 0.0 else None
        self.use_weighted_class_loss = config.WEIGHT_CLASS_LOSS
        self.heatmaps = heatmaps
        if joints_gt is not None:
            from .labels import SUPPORTED_METHODS
            if self.edge_label_method not in SUPPORTED_METHODS:
                raise NotImplementedError("EDGE_LABEL_METHOD=%r (4 and 6, 201 of the reference's 223 configs, are in scope)"
                                          % (self.edge_label_method,))
            if config.IMAGE_CENTRIC_SAMPLING:
                raise NotImplementedError("IMAGE_CENTRIC_SAMPLING is out of scope")
            if factor_list is None:
                raise ValueError("joints_gt needs factor_list (PoseEstimation.py:82-87)")
        self.mask_crowds = config.MASK_CROWDS
        self.detect_threshold = config.DETECT_THRESHOLD if config.DETECT_THRESHOLD <= 1.5 else None   # CG.py:28
        self.hybrid_k = config.HYBRID_K
        self.mpn_graph_type = config.GRAPH_TYPE
        self.normalize_node_distance = config.NORM_NODE_DISTANCE
        self.edge_features_to_use = config.EDGE_FEATURES_TO_USE
        self.pool_kernel_size = config.POOL_KERNEL_SIZE
        if self.mpn_graph_type not in ("knn", "fully"):
            raise NotImplementedError("GRAPH_TYPE=%r (knn and fully are in scope)" % (self.mpn_graph_type,))
        feats = set(self.edge_features_to_use)
        if not feats or not feats <= {"position", "connection_type"}:
            raise NotImplementedError("EDGE_FEATURES_TO_USE=%r (position / connection_type are in scope)"
                                      % (sorted(feats),))
        self._edge_feat_bits = ((nv.EDGE_FEAT_POSITION if "position" in feats else 0)
                                | (nv.EDGE_FEAT_TYPE if "connection_type" in feats else 0))
        if self.mask_crowds and self.masks is None:
            raise ValueError("MASK_CROWDS needs masks (PoseEstimation.py:73-74)")
        # capacities of the fixed-size device buffers (build-specific, optional config keys)
        top_k = self.hybrid_k if self.detect_threshold is not None else NO_THRESHOLD_K
        self._top_k = top_k
        self.max_det_per_type = int(getattr(config, "B200_MAX_DET_PER_TYPE", max(256, top_k)))
        self.cand_capacity = int(getattr(config, "B200_CAND_CAPACITY", max(4096, top_k)))
        max_nodes = int(getattr(config, "B200_MAX_NODES", min(2048, num_joints * self.max_det_per_type)))
        self.max_nodes = (max_nodes + 31) // 32 * 32
        # capacities the caller set explicitly are hard limits; the defaults grow on overflow (construct_graph)
        self._fixed_caps = {k for k in ("B200_MAX_DET_PER_TYPE", "B200_CAND_CAPACITY", "B200_MAX_NODES") if hasattr(config, k)}
        self.num_nodes_per_image = None
        self.num_edges_per_image = None

    def _grow(self, flags):
        """The fixed-size device buffers were too small for this input (the reference has no such limit): grow the ones
        that overflowed and tell the caller to run the detection again.  False when nothing can grow any further."""
        if flags & 8:
            return False
        ok = True
        if flags & 1:
            ok &= "B200_CAND_CAPACITY" not in self._fixed_caps and self.cand_capacity < 16384
            self.cand_capacity = min(16384, 4 * self.cand_capacity)
        if flags & 2:
            ok &= "B200_MAX_DET_PER_TYPE" not in self._fixed_caps and self.max_det_per_type < 4096
            self.max_det_per_type = min(4096, 4 * self.max_det_per_type)
        if flags & 4 or (flags & 2 and "B200_MAX_NODES" not in self._fixed_caps):
            ok &= "B200_MAX_NODES" not in self._fixed_caps and self.max_nodes < 16384
            self.max_nodes = min(16384, 2 * self.max_nodes)
        return bool(ok)

    def _launch_detect(self, stream=None):
        """Detection half (NMS -> candidates -> node layout -> kNN adjacency -> counts) on ``stream`` (default: the
        current one); the counts that size the outputs go to pinned host memory behind an event."""
        lib = nv.lib()
        dev = self.device
        sm_in = self.scoremaps
        if sm_in.dim() != 4 or sm_in.shape[1] != self.num_joints:
            raise ValueError("scoremaps must be [B, num_joints, H, W], got %s" % (tuple(sm_in.shape),))
        B, J, H, W = sm_in.shape
        fused = isinstance(sm_in, HeadStages)
        with torch.cuda.device(dev), torch.cuda.stream(stream if stream is not None else torch.cuda.current_stream(dev)):
            if fused:
                # the no-threshold path pads with zero-score pixels of the map (CG.py:1187): it needs the map in memory
                keep = sm_in.keep_scoremaps or self.detect_threshold is None
                sm = torch.empty(sm_in.shape, dtype=torch.float32, device=dev) if keep else None
                asm = sm_in.assembly(sm)
                sm_in.scoremaps = sm
                if stream is not None:
                    for pair in sm_in._keep:            # allocated on the caller's stream, read on this one
                        for t_ in pair:
                            t_.record_stream(stream)
            elif sm_in.device.type == "cpu":
                sm = torch.empty(sm_in.shape, dtype=torch.float32, device=dev)
                sm.copy_(sm_in, non_blocking=True)
            else:
                sm = nv.require_cuda(sm_in, "scoremaps").detach().float().contiguous()
                if stream is not None:
                    sm.record_stream(stream)        # allocated on the caller's stream, read on this one
            mask = self.masks.detach().float().contiguous() if self.mask_crowds else None
            p = nv.GcParams(
                batch=B, num_joints=J, height=H, width=W, pool_kernel=self.pool_kernel_size, top_k=self._top_k,
                use_threshold=int(self.detect_threshold is not None),
                threshold=float(self.detect_threshold if self.detect_threshold is not None else 0.0),
                graph_type=nv.GRAPH_FULLY if self.mpn_graph_type == "fully" else nv.GRAPH_KNN, knn_k=KNN_K,
                edge_features=self._edge_feat_bits,
                norm_factor=float(max(W, H) if self.normalize_node_distance else 1),     # CG.py:311-314
                cand_capacity=self.cand_capacity, max_det_per_type=self.max_det_per_type, max_nodes=self.max_nodes,
                scoremaps=sm.data_ptr() if sm is not None else None, mask=mask.data_ptr() if mask is not None else None)
            ws_bytes = int(lib.pgmp_gc_workspace_bytes(p))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            p.workspace, p.workspace_bytes = ws.data_ptr(), ws_bytes
            counts = torch.empty(2 + 2 * B + 1, dtype=torch.int64, device=dev)
            if fused:
                nv.check(lib.pgmp_gc_detect_fused(p, asm, counts.data_ptr(), nv.current_stream()))
                if sm is None:
                    sm = torch.empty((B, J, H, W), dtype=torch.float32, device="meta")   # shape only
            else:
                nv.check(lib.pgmp_gc_detect(p, counts.data_ptr(), nv.current_stream()))
            counts_h = torch.empty(counts.shape, dtype=torch.int64, pin_memory=True)
            counts_h.copy_(counts, non_blocking=True)
            event = torch.cuda.Event()
            event.record()
            used = torch.cuda.current_stream(dev)
        return dict(p=p, ws=ws, sm=sm, mask=mask, counts=counts, counts_h=counts_h, event=event, stream=used)

    def detect_async(self, stream=None):
        """Optional: start the detection half now, on ``stream`` (a side stream lets it -- and the host-to-device copy of
        pinned heatmaps -- overlap the previous batch's kernels).  ``construct_graph()`` then only waits for the counts,
        which have usually arrived long before, instead of draining the stream.  See ``pgmp_b200.pipeline``."""
        if self._pending is None:
            self._pending = self._launch_detect(stream)
        return self

    def construct_graph(self):
        lib = nv.lib()
        dev = self.device
        for attempt in range(6):
            d = self._pending or self._launch_detect(None)
            self._pending = None
            d["event"].synchronize()               # the one host wait of the graph constructor: the counts
            counts_h = d["counts_h"]
            flags = int(counts_h[-1])
            if not flags or not self._grow(flags):
                break
        if flags:
            raise RuntimeError("graph constructor capacity exceeded: " +
                               "; ".join(msg for bit, msg in nv.GC_FLAGS.items() if flags & bit))
        p, ws, sm = d["p"], d["ws"], d["sm"]
        B, J, H, W = sm.shape
        with torch.cuda.device(dev):
            stream = nv.current_stream()
            cur = torch.cuda.current_stream(dev)
            if d["stream"] != cur:                 # the emit kernels read the detection's workspace on this stream
                cur.wait_event(d["event"])
                for t_ in (ws, sm, d["counts"]) + ((d["mask"],) if d["mask"] is not None else ()):
                    if t_.device.type == "cuda":
                        t_.record_stream(cur)
            N, E = int(counts_h[0]), int(counts_h[1])
            self.num_nodes_per_image = counts_h[2:2 + B].clone()
            self.num_edges_per_image = counts_h[2 + B:2 + 2 * B].clone()

            feat = self.features
            C_ = feat.shape[1] if feat is not None else 0
            F_ = (2 if self._edge_feat_bits & nv.EDGE_FEAT_POSITION else 0) + \
                 (J if self._edge_feat_bits & nv.EDGE_FEAT_TYPE else 0)
            x = torch.empty((N, C_), dtype=torch.float32, device=dev)
            edge_attr = torch.empty((E, F_), dtype=torch.float32, device=dev)
            edge_index = torch.empty((2, E), dtype=torch.int64, device=dev)
            joint_det = torch.empty((N, 3), dtype=torch.int64, device=dev)
            joint_scores = torch.empty((N,), dtype=torch.float32, device=dev)
            batch_index = torch.empty((N,), dtype=torch.int64, device=dev)
            tags = self.tagmaps
            tag_dim, joint_tags, tags_c = 1, None, None
            stage_tags = isinstance(tags, HeadStages)
            if stage_tags:
                s1t = tags.terms[0][0].detach().contiguous()
                if s1t.shape[1] != 2 * J or tuple(tags.shape[2:]) != (H, W):
                    raise ValueError("HeadStages as tagmaps: scoremap_1 must hold J heatmaps + J tag maps for %d x %d maps" % (H, W))
                joint_tags = torch.empty((N,), dtype=torch.float32, device=dev)
            elif tags is not None:
                if tags.dim() not in (4, 5):
                    raise ValueError("tagmaps must be [B,J,H,W] or [B,J,H,W,T]")
                tag_dim = tags.shape[4] if tags.dim() == 5 else 1
                tags_c = tags.detach().float().contiguous()
                joint_tags = torch.empty((N,) if tags.dim() == 4 else (N, tag_dim), dtype=torch.float32, device=dev)
            o = nv.GcOutputs(total_nodes=N, total_edges=E)
            fused = isinstance(feat, ConvUpsampleFeatures)
            if fused:
                if feat.size != (H, W):
                    raise ValueError("ConvUpsampleFeatures size %s != heatmap size %s" % (feat.size, (H, W)))
            elif feat is not None:
                if feat.dtype != torch.float32:
                    raise TypeError("features must be float32 (the reference gathers fp32, CG.py:265)")
                fd = feat.detach()
                o.features = fd.data_ptr()
                o.feat_stride_b, o.feat_stride_c, o.feat_stride_y, o.feat_stride_x = fd.stride()
                o.channels = C_
                o.x = x.data_ptr()
            if tags_c is not None:
                o.tagmaps, o.tag_dim, o.joint_tags = tags_c.data_ptr(), tag_dim, joint_tags.data_ptr()
            o.edge_attr, o.edge_index = edge_attr.data_ptr(), edge_index.data_ptr()
            o.joint_det, o.joint_scores, o.batch_index = joint_det.data_ptr(), joint_scores.data_ptr(), batch_index.data_ptr()
            nv.check(lib.pgmp_gc_emit(p, o, stream))
            if stage_tags:        # up(scoremap_1)[:, J:] at the detections only: the tag maps are never up-sampled
                nv.check(lib.pgmp_gc_gather_stage_tags(s1t.data_ptr(), s1t.shape[1], J, s1t.shape[2], s1t.shape[3], H, W,
                                                       joint_det.data_ptr(), batch_index.data_ptr(), N, joint_tags.data_ptr(),
                                                       stream))
            if fused and N > 0:
                wt, bias = feat.pack(dev)
                fm = feat.feat.detach()
                gp = nv.GatherConvParams(
                    features=fm.data_ptr(), feat_stride_b=fm.stride(0), feat_stride_c=fm.stride(1),
                    feat_stride_y=fm.stride(2), feat_stride_x=fm.stride(3), cin=fm.shape[1], height=fm.shape[2],
                    width=fm.shape[3], cout=C_, out_height=H, out_width=W, weight_t=wt.data_ptr(), bias=bias.data_ptr(),
                    joint_det=joint_det.data_ptr(), batch_index=batch_index.data_ptr(), num_nodes=N, x=x.data_ptr())
                nv.check(lib.pgmp_gc_gather_conv(gp, stream))
            ws.record_stream(torch.cuda.current_stream())
        if feat is not None and not fused and feat.requires_grad and torch.is_grad_enabled():
            x = _GatherNodeFeatures.apply(feat, x, batch_index, joint_det)
        if fused and N > 0 and torch.is_grad_enabled() and (
                feat.feat.requires_grad or feat.weight.requires_grad or (feat.bias is not None and feat.bias.requires_grad)):
            x = _GatherConvFeatures.apply(feat.feat, feat.weight, feat.bias, x, batch_index, joint_det, (H, W))
        if self.joints_gt is None:
            # the reference's 15-tuple (ConstructGraph.py:248-249); label slots are None at inference (:243-246)
            return (x, edge_attr, edge_index, None, None, None, None, joint_det, None, None, None, joint_scores,
                    batch_index, None, joint_tags)
        from . import labels as L
        lab = L.build_labels(self, joint_det, edge_index, batch_index, self.num_nodes_per_image.tolist())
        if self.node_dropout is not None and not self.testing:
            (x, edge_attr, edge_index, joint_det, joint_scores, batch_index, joint_tags, lab,
             nodes_per_image) = L.node_dropout(self.node_dropout, x, edge_attr, edge_index, joint_det, joint_scores,
                                               batch_index, joint_tags, lab, B)
            self.num_nodes_per_image = nodes_per_image.cpu()
            self.num_edges_per_image = torch.bincount(batch_index[edge_index[0]], minlength=B).cpu()
        if self.use_weighted_class_loss and lab["class_mask"] is not None:      # ConstructGraph.py:170-176, indices as written there
            hm = self.heatmaps.to(dev)
            wts = hm[batch_index, lab["node_classes"], joint_det[:, 1], joint_det[:, 2]]
            lab["class_mask"] = torch.where(wts < 0.1, torch.full_like(wts, 0.1), wts) * lab["class_mask"]
        return (x, edge_attr, edge_index, lab["edge_labels"], lab["node_labels"], lab["node_classes"], None, joint_det,
                lab["label_mask"], lab["label_mask_node"], lab["class_mask"], joint_scores, batch_index,
                lab["node_persons"], joint_tags)


def hr_process_output(output, mode, num_joints):
    """Drop-in for the closure ``create_process_func_hr(config)`` returns (src/Models/HigherHRNet/hrnet.py:587-611):
    ``output = ((scoremap_1, scoremap_2), features)`` -> ``(scoremaps, features, tags)``.  The bilinear up-sampling of the
    half-resolution stage, the average with the full-resolution stage and the tag maps come out of ONE CUDA kernel
    (``pgmp_gc_assemble_scoremaps``) instead of interpolate + add + divide over the full maps."""
    (s1, s2), features = output
    if mode == "large":
        return s2, features, s1[:, num_joints:]
    if mode not in ("avg", "small"):
        raise NotImplementedError(mode)
    nv.require_cuda(s1, "scoremap_1", torch.float32)
    nv.require_cuda(s2, "scoremap_2", torch.float32)
    s1c, s2c = s1.detach().contiguous(), s2.detach().contiguous()
    B, C1, h, w = s1c.shape
    H, W = s2c.shape[2], s2c.shape[3]
    if C1 < num_joints or s2c.shape[0] != B or (mode == "avg" and s2c.shape[1] != num_joints):
        raise ValueError("scoremap_1 must be [B, >= J, h, w] and scoremap_2 [B, J, H, W]")
    scoremaps = torch.empty((B, num_joints, H, W), dtype=torch.float32, device=s1.device)
    tags = torch.empty((B, C1 - num_joints, H, W), dtype=torch.float32, device=s1.device)
    with torch.cuda.device(s1.device):
        nv.check(nv.lib().pgmp_gc_assemble_scoremaps(s1c.data_ptr(), s2c.data_ptr(), B, C1, num_joints, h, w, H, W,
                                                     0 if mode == "avg" else 1, scoremaps.data_ptr(),
                                                     tags.data_ptr() if C1 > num_joints else None, nv.current_stream()))
    return scoremaps, features, tags


def get_graph_constructor(config, **kwargs):
    """src/graph_constructor/__init__.py:4-5."""
    return NaiveGraphConstructor(config=config, **kwargs)
