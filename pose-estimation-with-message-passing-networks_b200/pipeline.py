"""Software-pipelined grouping path over a sequence of batches.

The reference runs ``construct_graph()`` then ``mpn.forward()`` per batch (``PoseEstimation.py:82-93, 232-243``).  The
output sizes of the graph constructor are data dependent, so one host read (the node / edge counts) sits between its
detection half and everything after it; issued back to back, that read drains the stream once per batch and the ~60
short launches that follow it start on an idle GPU.  ``GroupingPipeline`` keeps the drop-in API and hides the wait:
the detection half of batch ``i + 1`` (NMS, candidates, kNN adjacency, counts) -- and, for pinned HOST heatmaps, its
host-to-device copy -- is launched on a side stream before batch ``i`` is finished, so by the time
``construct_graph()`` needs the counts they have long arrived and the GPU never runs dry.  Consecutive batches run
their emit half and the network on two alternating compute streams (``interleave=True``): a batch is a dependent chain
of ~45 kernels with a drain / fill bubble at every boundary, and the neighbouring batch's kernels fill those bubbles
(measured: 3.50 -> 3.19 ms per 32-image batch); the caller's stream waits for a batch before it is handed out, so the
results are used exactly as after the serial calls.  Every batch still does the full work; results are identical to the
serial calls (same kernels, same order per batch).

    pipe = GroupingPipeline(gc_config, mpn, num_joints, device)
    for graph, (preds_edge, preds_node, preds_class) in pipe.run(batches):   # batches: dicts of construct_graph kwargs
        ...

With ``group=dict(node_threshold=..., cc_method=...)`` the grouping tail (sigmoid / threshold / GAEC / persons,
``Utils.group_persons``) of batch ``i`` runs on a third stream while batch ``i + 1`` goes through the network -- the greedy
contraction is a sequential algorithm that keeps one CTA per image busy for milliseconds on a fraction of the SMs -- and
``run`` yields ``(graph, preds, groups)``.
"""

import torch

from .graph_constructor import get_graph_constructor


class GroupingPipeline:
    def __init__(self, gc_config, model, num_joints, device, testing=True, group=None, interleave=True):
        self.gc_config, self.model, self.num_joints = gc_config, model, num_joints
        self.device = torch.device(device)
        self.testing = testing
        self.side = torch.cuda.Stream(device=self.device)
        # consecutive batches go through emit + network on two alternating streams: the ~45 kernels of a batch are a
        # dependent chain with a drain / fill bubble at every boundary, and the neighbouring batch's kernels fill them
        self.compute = [torch.cuda.Stream(device=self.device) for _ in range(2)] if interleave else None
        self._count = 0
        self._repacked = None                               # event after a batch that (re)built the model's packed weights
        self.group = dict(group) if group is not None else None
        self.group_stream = torch.cuda.Stream(device=self.device) if group is not None else None

    def _submit(self, batch):
        """Start the detection half of ``batch`` on the side stream; returns the graph constructor holding it."""
        main = torch.cuda.current_stream(self.device)
        gc = get_graph_constructor(self.gc_config, scoremaps=batch["scoremaps"], tagmaps=batch.get("tagmaps"),
                                   features=batch.get("features"), joints_gt=None, factor_list=None,
                                   masks=batch.get("masks"), device=self.device, testing=self.testing, heatmaps=None,
                                   num_joints=self.num_joints)
        ev = torch.cuda.Event()                              # the inputs were produced on the main stream (the backbone)
        ev.record(main)
        gc._inputs_ready = ev
        if batch["scoremaps"].device.type == "cuda":
            self.side.wait_event(ev)
        gc.detect_async(self.side)
        return gc

    def _finish_interleaved(self, gc):
        """``_finish`` on one of the two compute streams; the caller's stream waits for the batch before it is handed out."""
        main = torch.cuda.current_stream(self.device)
        cs = self.compute[self._count & 1]
        self._count += 1
        cs.wait_event(gc._inputs_ready)
        if self._repacked is not None:                       # the other stream built the packed weights this batch reads
            cs.wait_event(self._repacked)
            self._repacked = None
        cache = getattr(self.model, "_pack_cache", None)
        with torch.cuda.stream(cs):
            out = self._finish(gc)
            done = torch.cuda.Event()
            done.record(cs)
        if getattr(self.model, "_pack_cache", None) is not cache:
            self._repacked = done
        main.wait_event(done)
        for t in list(out[0]) + [x for lst in out[1] for x in lst]:
            if torch.is_tensor(t) and t.device.type == "cuda":
                t.record_stream(main)                        # allocated on the compute stream, consumed on the caller's
        return out

    def _finish(self, gc):
        ret = gc.construct_graph()
        with torch.no_grad():
            pe, pn, pc, _ = self.model(ret[0], ret[1], ret[2], node_labels=None, edge_labels=None, batch_index=ret[12],
                                       node_mask=None, node_types=ret[7][:, 2])
        if self.group is None:
            return ret, (pe, pn, pc)
        from .Utils import group_persons_async
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.group_stream.wait_event(ev)                   # the logits are ready on the main stream
        pend = group_persons_async(ret[7], pn[-1], ret[2], pe[-1], pc[-1], ret[12], self.num_joints,
                                   detector_scores=ret[11], nodes_per_image=gc.num_nodes_per_image,
                                   edges_per_image=gc.num_edges_per_image, stream=self.group_stream, **self.group)
        return ret, (pe, pn, pc), pend

    def run(self, batches):
        """Generator over ``(graph 15-tuple, (preds_edge, preds_node, preds_class))`` -- plus the groups of
        ``Utils.group_persons`` when the pipeline was built with ``group=`` -- one per batch, in order."""
        prev, waiting = None, None                         # constructor with its detection in flight; batch whose groups are in flight
        for batch in batches:
            cur = self._submit(batch)
            if prev is not None:
                out = self._finish_interleaved(prev) if self.compute is not None else self._finish(prev)
                if self.group is None:
                    yield out
                else:
                    if waiting is not None:
                        yield waiting[0], waiting[1], waiting[2].result()
                    waiting = out
            prev = cur
        if prev is not None:
            out = self._finish_interleaved(prev) if self.compute is not None else self._finish(prev)
            if self.group is None:
                yield out
            else:
                if waiting is not None:
                    yield waiting[0], waiting[1], waiting[2].result()
                yield out[0], out[1], out[2].result()
        elif waiting is not None:
            yield waiting[0], waiting[1], waiting[2].result()
