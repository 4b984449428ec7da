"""TEST INFRASTRUCTURE.  Mints tests/golden/vq_decode_{llamagen,chameleon}.json from the UNMODIFIED reference modules
(run in the build container, where /root/reference exists):

  llamagen   llamagen/tokenizer/tokenizer_image/vq_model.py: VectorQuantizer.get_codebook_entry (L2-normalised codebook)
             -> post_quant_conv -> Decoder, i.e. VQModel.decode_code (:52-55) on a narrow instance of the same classes
  chameleon  lumina_mgpt/model/chameleon_vae_ori/vqgan.py: VQModel(ddconfig).quantize.get_codebook_entry -> decode
             (what image_tokenizer.py:116-121 calls)

    python oracle/mint_vq_golden.py
"""
import importlib.util
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.vq_case import fill_state  # noqa: E402

REF = Path("/root/reference")


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def mint_llamagen():
    vq = load(REF / "llamagen/tokenizer/tokenizer_image/vq_model.py", "ref_llamagen_vq")
    torch.manual_seed(0)
    z, e_dim, n_e = 32, 8, 96
    m = torch.nn.Module()
    m.decoder = vq.Decoder(z_channels=z, ch=32, ch_mult=(1, 2, 2), num_res_blocks=1)
    m.quantize = vq.VectorQuantizer(n_e, e_dim, 0.25, 0.0, True, False)
    m.post_quant_conv = torch.nn.Conv2d(e_dim, z, 1)
    m.eval()
    shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(fill_state(shapes, 11))
    g = torch.Generator().manual_seed(5)
    B, h, w = 2, 4, 6
    codes = torch.randint(0, n_e, (B * h * w,), generator=g)
    with torch.no_grad():
        quant = m.quantize.get_codebook_entry(codes, (B, e_dim, h, w), True)
        px = m.decoder(m.post_quant_conv(quant))
    return dict(family="llamagen", seed=11, shapes=shapes, codes=codes.tolist(), batch=B, h=h, w=w, l2_norm=True,
                out_shape=list(px.shape), pixels=[round(float(x), 7) for x in px.flatten().tolist()])


def mint_chameleon():
    vq = load(REF / "lumina_mgpt/model/chameleon_vae_ori/vqgan.py", "ref_chameleon_vq")
    torch.manual_seed(0)
    dd = dict(double_z=False, z_channels=32, resolution=32, in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2, 2],
              num_res_blocks=1, attn_resolutions=[16], dropout=0.0)
    m = vq.VQModel(ddconfig=dd, n_embed=80, embed_dim=16)
    m.eval()
    shapes = {k: list(v.shape) for k, v in m.state_dict().items()
              if k.startswith(("decoder.", "post_quant_conv.", "quantize.embedding."))}
    full = m.state_dict()
    full.update(fill_state(shapes, 12))
    m.load_state_dict(full)
    g = torch.Generator().manual_seed(6)
    h, w = 8, 8
    codes = torch.randint(0, 80, (h * w,), generator=g)
    with torch.no_grad():
        entry = m.quantize.get_codebook_entry(codes, (1, h, w, 16))
        px = m.decode(entry)
    return dict(family="chameleon", seed=12, shapes=shapes, codes=codes.tolist(), batch=1, h=h, w=w, l2_norm=False,
                out_shape=list(px.shape), pixels=[round(float(x), 7) for x in px.flatten().tolist()])


def mint_emu3():
    """emu3/tokenizer/modeling_emu3visionvq.py: Emu3VisionVQModel(config).decode(codes [B, h, w]) on a narrow config."""
    import types
    pkg = types.ModuleType("ref_emu3_tok")
    pkg.__path__ = [str(REF / "emu3" / "tokenizer")]
    sys.modules["ref_emu3_tok"] = pkg
    cfgm = load(REF / "emu3/tokenizer/configuration_emu3visionvq.py", "ref_emu3_tok.configuration_emu3visionvq")
    sys.modules["ref_emu3_tok.configuration_emu3visionvq"] = cfgm
    spec = importlib.util.spec_from_file_location("ref_emu3_tok.modeling_emu3visionvq",
                                                  REF / "emu3/tokenizer/modeling_emu3visionvq.py")
    vq = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = vq
    spec.loader.exec_module(vq)
    torch.manual_seed(0)
    cfg = cfgm.Emu3VisionVQConfig(codebook_size=72, embed_dim=4, z_channels=4, ch=32, ch_mult=[1, 2, 2], num_res_blocks=1,
                                  attn_resolutions=[2], temporal_downsample_factor=4)
    m = vq.Emu3VisionVQModel(cfg)
    m.eval()
    full = m.state_dict()
    shapes = {k: list(v.shape) for k, v in full.items()
              if k.startswith(("decoder.", "post_quant_conv.", "quantize.embedding."))}
    full.update(fill_state(shapes, 13))
    m.load_state_dict(full)
    g = torch.Generator().manual_seed(7)
    B, h, w = 2, 5, 4
    codes = torch.randint(0, 72, (B, h, w), generator=g)
    with torch.no_grad():
        px = m.decode(codes)
    return dict(family="emu3", seed=13, shapes=shapes, codes=codes.flatten().tolist(), batch=B, h=h, w=w, l2_norm=False,
                out_shape=list(px.shape), pixels=[round(float(x), 7) for x in px.flatten().tolist()])


if __name__ == "__main__":
    for fn in (mint_llamagen, mint_chameleon, mint_emu3):
        g = fn()
        p = ROOT / "tests" / "golden" / f"vq_decode_{g['family']}.json"
        p.write_text(json.dumps(g))
        print(p, p.stat().st_size, "bytes;", len(g["shapes"]), "tensors; pixels", g["out_shape"])
