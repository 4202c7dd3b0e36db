"""Generate the committed golden vectors under tests/golden/*.npz FROM THE ORACLE.

    python tests/golden/make_golden.py            # rewrites the .npz files

torch_geometric is not installable here (oracle/__init__.py), so these fixtures pin the
oracle against regressions and give the GPU box a reference that needs nothing but numpy —
they are not outputs of PyG itself.  Inputs are stored next to outputs; tests recompute the
outputs from the STORED inputs, so RNG-stream changes cannot invalidate them.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from oracle import (EncoderRef, GraphNormRef, RGCNConvRef, degree_ref, edge_type_bucket_ref,  # noqa: E402
                    rel_csr_ref, soft_masking_ref, transposed_csr_ref)

HERE = Path(__file__).resolve().parent

CASES = {
    # name: (num_nodes, num_edges, in_feat, hidden, out_dim, seed)
    "cornell_shaped": (183, 300, 96, 8, 12, 42),      # Cornell-shaped topology, narrow features (file size)
    "multi_edge_self_loop": (12, 80, 10, 4, 6, 7),
    "hub": (60, 900, 16, 4, 6, 9),
}


def make_inputs(name):
    n, e, fin, hidden, out_dim, seed = CASES[name]
    g = torch.Generator().manual_seed(seed)
    if name == "hub":
        dst = torch.where(torch.rand(e, generator=g) < 0.5, torch.zeros(e, dtype=torch.long),
                          torch.randint(0, n, (e,), generator=g))
        ei = torch.stack([torch.randint(0, n, (e,), generator=g), dst])
    else:
        ei = torch.randint(0, n, (2, e), generator=g)
    x = torch.randn(n, fin, generator=g)
    mask = torch.rand(n, generator=g) < 0.3
    token = torch.randn(1, fin, generator=g) * 0.1
    torch.manual_seed(seed)
    enc = EncoderRef(fin, hidden, out_dim, dropout_rate=0.0, use_checkpoint=False)
    with torch.no_grad():
        for k in range(1, 5):
            getattr(enc, f"gnorm{k}").mean_scale.uniform_(0.5, 1.0)
            getattr(enc, f"rgcn{k}").bias.uniform_(-0.1, 0.1)
    inputs = {"edge_index": ei.numpy(), "x": x.numpy(), "mask": mask.numpy(), "token": token.numpy(),
              "dims": np.array([n, e, fin, hidden, out_dim], dtype=np.int64)}
    for k, v in enc.state_dict().items():
        inputs["param/" + k] = v.numpy()
    return inputs


def compute_outputs(inputs):
    n, e, fin, hidden, out_dim = (int(v) for v in inputs["dims"])
    ei = torch.from_numpy(inputs["edge_index"])
    x = torch.from_numpy(inputs["x"]).double()
    mask = torch.from_numpy(inputs["mask"])
    token = torch.from_numpy(inputs["token"]).double()
    enc = EncoderRef(fin, hidden, out_dim, dropout_rate=0.0, use_checkpoint=False).double()
    enc.load_state_dict({k[len("param/"):]: torch.from_numpy(v).double() for k, v in inputs.items()
                         if k.startswith("param/")})
    et = edge_type_bucket_ref(ei, n)
    live = sorted(set(et.tolist())) or [0]
    slot_of = {r: s for s, r in enumerate(live)}
    slot = np.array([slot_of[int(t)] for t in et.tolist()], dtype=np.int64)
    rowptr, col, perm = rel_csr_ref(ei[0].numpy(), ei[1].numpy(), slot, n, len(live))
    rowptr_t, seg_t, w_t, perm_t = transposed_csr_ref(ei[0].numpy(), ei[1].numpy(), slot, n, len(live))
    x_masked = soft_masking_ref(x, mask, token, 0.7)
    x_masked_f32 = soft_masking_ref(torch.from_numpy(inputs["x"]), mask, torch.from_numpy(inputs["token"]), 0.7)
    conv1 = enc.rgcn1(x, ei, et)
    norm1 = enc.gnorm1(conv1)
    fused, layers = enc(x, ei, return_layers=True)
    fused_masked = enc(x_masked, ei)
    out = {
        "degree_src": degree_ref(ei[0], n).numpy(), "edge_type": et.numpy(),
        "live_rels": np.array(live, dtype=np.int64),
        "rowptr": rowptr, "col": col, "perm": perm,
        "rowptr_t": rowptr_t, "seg_t": seg_t, "w_t": w_t, "perm_t": perm_t,
        "x_masked_f32": x_masked_f32.numpy(),
        "conv1": conv1.detach().numpy(), "norm1": norm1.detach().numpy(),
        "fused": fused.detach().numpy(), "fused_masked": fused_masked.detach().numpy(),
    }
    for i, l in enumerate(layers):
        out[f"layer{i+1}"] = l.detach().numpy()
    return out


def build_cases():
    """name -> {output name: array}, recomputed from the STORED inputs when they exist."""
    cases = {}
    for name in CASES:
        path = HERE / f"{name}.npz"
        if path.exists():
            stored = np.load(path)
            inputs = {k: stored[k] for k in stored.files if k in ("edge_index", "x", "mask", "token", "dims")
                      or k.startswith("param/")}
        else:
            inputs = make_inputs(name)
        cases[name] = compute_outputs(inputs)
    return cases


def main():
    for name in CASES:
        inputs = make_inputs(name)
        outputs = compute_outputs(inputs)
        np.savez_compressed(HERE / f"{name}.npz", **inputs, **outputs)
        size = (HERE / f"{name}.npz").stat().st_size
        print(f"wrote {name}.npz ({size/1024:.1f} KiB)")


if __name__ == "__main__":
    main()
