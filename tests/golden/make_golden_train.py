"""Training-mode golden vectors of the reference MPN (round-2 groundwork).

    python tests/golden/make_golden_train.py

Runs the UNMODIFIED reference ``NodeClassificationMPNSimple`` in ``train()`` mode (BatchNorm batch statistics) under
torch autograd, with float64 tensors, on a small graph and stores, per case: the training-mode logits, the scalar loss
``L = sum_k <c_k, pred_k>`` (fixed random ``c``), the gradient w.r.t. the node input ``x`` and, per parameter, the
gradient's L2 norm plus 64 sampled entries (fixed indices), and the updated BatchNorm running statistics.
Needs /root/reference; the fixtures travel.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import pgmp_b200  # noqa: E402
import pgmp_b200.synthetic as synthetic  # noqa: E402
import ref_shims  # noqa: E402
from cases import GC_CASES, TRAIN_CASES, gc_config_for, mpn_config_for, train_loss_weights, sample_indices  # noqa: E402
from make_golden import run_reference_gc  # noqa: E402


def run(mpn, cfg, g, seed, dtype):
    model = mpn.NodeClassificationMPNSimple(cfg)
    synthetic.synth_mpn_state_dict(model, seed)            # float32 weights, exactly those of the tests
    model = model.to(dtype).train()
    # float64 run: the reference allocates two scratch buffers with an explicit dtype=torch.float32
    # (TypeAwareNodeUpdate.forward, layers.py:270); the harness lets those follow the run's dtype, nothing else changes
    zeros = torch.zeros
    if dtype == torch.float64:
        torch.zeros = lambda *a, **k: zeros(*a, **{**k, "dtype": dtype if k.get("dtype") == torch.float32 else k.get("dtype", None)})
    try:
        x = torch.from_numpy(g["x"]).to(dtype).clone().requires_grad_(True)
        pe, pn, pc, _ = model(x, torch.from_numpy(g["edge_attr"]).to(dtype), torch.from_numpy(g["edge_index"]),
                              node_types=torch.from_numpy(g["joint_det"][:, 2]))
        preds = list(pe) + list(pn[:-1]) + list(pc[:-1])      # the trailing node / class entries repeat the last step (:93-94)
        cw = train_loss_weights([tuple(p.shape) for p in preds], seed)
        loss = sum((p * torch.from_numpy(c).to(dtype)).sum() for p, c in zip(preds, cw))
        loss.backward()
    finally:
        torch.zeros = zeros
    return model, x, pe, pn, pc, loss


def main():
    cg, mpn = ref_shims.load_reference()
    for name, (gc_name, maker, over, seed) in TRAIN_CASES.items():
        g = run_reference_gc(cg, gc_name)
        cfg = mpn_config_for(pgmp_b200.config, maker, over)
        # float64 tensors through the unmodified reference code: the exact gradient.  The same run in float32 (what
        # train.py does) deviates from it by the amount recorded in fp32_* -- its own round-off, amplified by the
        # BatchNorm statistics over ~20 k edges -- which is the tolerance a float32 implementation can be held to.
        model, x, pe, pn, pc, loss = run(mpn, cfg, g, seed, torch.float64)
        _, x32, _, _, _, loss32 = run(mpn, cfg, g, seed, torch.float32)
        gx, gx32 = x.grad.numpy(), x32.grad.double().numpy()
        out = {"loss": np.float64(loss.item()), "grad_x": gx.astype(np.float32), "n_edge": np.int64(len(pe)), "n_node": np.int64(len(pn)),
               "fp32_grad_x_l2rel": np.float64(np.linalg.norm(gx32 - gx) / np.linalg.norm(gx)),
               "fp32_grad_x_maxrel": np.float64(np.abs(gx32 - gx).max() / np.abs(gx).max()),
               "fp32_loss": np.float64(loss32.item())}
        for i, a in enumerate(pe):
            out[f"edge_{i}"] = a.detach().numpy().astype(np.float32)
        for i, a in enumerate(pn):
            out[f"node_{i}"] = a.detach().numpy().astype(np.float32)
        for i, a in enumerate(pc):
            out[f"class_{i}"] = a.detach().numpy().astype(np.float32)
        for pname, p in model.named_parameters():
            gr = p.grad.numpy().ravel() if p.grad is not None else np.zeros(p.numel())
            out["gnorm/" + pname] = np.float64(np.linalg.norm(gr))
            out["gsamp/" + pname] = gr[sample_indices(gr.size, pname)]
        for bname, b in model.named_buffers():
            if bname.endswith("running_mean") or bname.endswith("running_var"):
                out["buf/" + bname] = b.numpy()
        np.savez_compressed(os.path.join(HERE, f"train_{name}.npz"), **out)
        print(f"train_{name}: loss {loss.item():.6f} |grad_x| {np.linalg.norm(gx):.4f}; float32 run: grad_x l2-rel "
              f"{out['fp32_grad_x_l2rel']:.2e} max-rel {out['fp32_grad_x_maxrel']:.2e}")


if __name__ == "__main__":
    main()
