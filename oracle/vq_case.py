"""TEST INFRASTRUCTURE (not shipped, not on the product path).  Deterministic weights for the VQ-decoder goldens: the
minting script (oracle/mint_vq_golden.py) builds the reference's modules, overwrites every tensor of their state dict from
`fill_state`, runs the reference and records {key: shape}, the codes and the pixels; the GPU test regenerates the same
tensors from the recorded shapes and must reproduce the pixels with sjd_b200.vq_decode.VQDecoder."""
import torch


def fill_state(shapes: dict, seed: int) -> dict:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros(shp, dtype=torch.long)
            continue
        t = torch.randn(shp, generator=g)
        if k.endswith("running_var"):
            t = 0.5 + t.abs()
        elif k.endswith("running_mean"):
            t = 0.1 * t
        elif k.endswith("norm_layer.weight"):
            t = 1.0 + 0.2 * t
        elif k.endswith("norm.weight") or ".norm1.weight" in k or ".norm2.weight" in k or k.endswith("norm_out.weight"):
            t = 1.0 + 0.2 * t                                  # GroupNorm scales around one
        elif k.endswith(".bias"):
            t = 0.1 * t
        elif k == "quantize.embedding.weight":
            t = 0.5 * t
        else:                                                   # convolutions: keep activations O(1) through the stack
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            t = t / (fan_in ** 0.5)
        sd[k] = t
    return sd
