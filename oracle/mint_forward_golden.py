"""ORACLE tooling (build container only, needs /root/reference): pin oracle/ref_forward.RefStack against the
reference's OWN model code run in fp32 on CPU.

  python oracle/mint_forward_golden.py        # writes tests/golden/forward_llamagen.npz, forward_chameleon.npz

  * LlamaGen:  llamagen/llamagen.py Transformer (GPT, c2i) with its static KV cache — cond-token prefill, a window, a
               rolled-back window (llamagen.py:368-405, :222-278, :441-467).
  * Chameleon: lumina_mgpt/model/chameleon/modeling_chameleon.py ChameleonForConditionalGeneration (the Lumina-mGPT
               backbone: QK-LayerNorm, rotate-half RoPE) driven with the reference's renewed _update_causal_mask
               (scheduler/jacobi_iteration_lumina_mgpt.py:1256-1336) and 3-D Jacobi window masks, CFG-uncond row with
               its prompt prefix hidden (:742-770).
The module weights are oracle.ref_forward.random_weights(seed) copied INTO the reference modules, so a fixture only
stores the seed, the call sequence and the reference logits.
tests/test_oracle_golden.py replays the calls through RefStack(emulate_bf16=False) and compares (<= 2e-4 abs).
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("SJD_REFERENCE", "/root/reference"))
sys.path.insert(0, str(REPO))
sys.dont_write_bytecode = True
OUT = REPO / "tests" / "golden"
LLAMAGEN_SEED, CHAMELEON_SEED = 101, 202


def pack_common(layers, embed, final_norm, lm_head):
    return {"embed": embed, "final_norm": final_norm, "lm_head": lm_head, "layers": layers}


def flat(w: dict) -> dict:
    out = {"embed": w["embed"].numpy(), "final_norm": w["final_norm"].numpy(), "lm_head": w["lm_head"].numpy()}
    for i, L in enumerate(w["layers"]):
        for k, v in L.items():
            out[f"layers.{i}.{k}"] = v.numpy()
    return out


@torch.no_grad()
def mint_llamagen():
    sys.path.insert(0, str(REF))
    from llamagen.llamagen import ModelArgs, Transformer
    torch.manual_seed(11)
    args = ModelArgs(dim=256, n_layer=2, n_head=4, vocab_size=1024, block_size=64, cls_token_num=1, num_classes=10,
                     model_type="c2i", class_dropout_prob=0.1)
    m = Transformer(args).float().eval()
    # weights come from oracle.ref_forward.random_weights(seed) so that the test can regenerate them instead of the
    # fixture carrying megabytes of parameters (the reference zero-inits the head, llamagen.py:342)
    from oracle import ref_forward as RF
    ff = m.layers[0].feed_forward.w1.weight.shape[0]
    rcfg = RF.StackConfig(args.n_layer, args.dim, args.n_head, args.n_head, args.dim // args.n_head, ff, args.vocab_size,
                          args.norm_eps, rope_interleaved=True, family="llamagen")
    w = RF.random_weights(rcfg, seed=LLAMAGEN_SEED, std=0.05)
    m.tok_embeddings.weight.copy_(w["embed"]); m.norm.weight.copy_(w["final_norm"]); m.output.weight.copy_(w["lm_head"])
    for L, wl in zip(m.layers, w["layers"]):
        L.attention_norm.weight.copy_(wl["attn_norm"]); L.attention.wqkv.weight.copy_(wl["wqkv"])
        L.attention.wo.weight.copy_(wl["wo"]); L.ffn_norm.weight.copy_(wl["ffn_norm"])
        L.feed_forward.w1.weight.copy_(wl["w_gate_up"][:ff]); L.feed_forward.w3.weight.copy_(wl["w_gate_up"][ff:])
        L.feed_forward.w2.weight.copy_(wl["w_down"])
    rows, max_len = 2, 40
    m.setup_caches(max_batch_size=rows, max_seq_length=max_len, dtype=torch.float32)
    g = torch.Generator().manual_seed(3)
    calls, logits = [], []
    cond = torch.tensor([3, 10])   # class 3, and the CFG "null" class (= num_classes)
    lg, _ = m(None, cond, input_pos=torch.arange(0, 1))
    cond_embeds = m.cls_embedding(cond, train=False)[:, :1]
    calls.append(("embeds", 0, 1))
    logits.append(lg.numpy())
    ids_all = []
    for kv_len, W in ((1, 6), (4, 9), (13, 1)):    # second call re-writes slots 4..6 (a Jacobi roll-back)
        ids = torch.randint(0, args.vocab_size, (rows, W), generator=g)
        lg, _ = m(ids, None, input_pos=torch.arange(kv_len, kv_len + W))
        calls.append(("ids", kv_len, W))
        ids_all.append(ids.numpy())
        logits.append(lg.numpy())
    np.savez_compressed(OUT / "forward_llamagen.npz", cond_embeds=cond_embeds.numpy(), seed=np.array([LLAMAGEN_SEED]),
                        calls=np.array([(0 if c[0] == "embeds" else 1, c[1], c[2]) for c in calls]),
                        **{f"ids{i}": a for i, a in enumerate(ids_all)}, **{f"logits{i}": a for i, a in enumerate(logits)},
                        cfg=np.array([args.n_layer, args.dim, args.n_head, args.n_head, args.dim // args.n_head, ff,
                                      args.vocab_size]),
                        grid=np.array([8]), rope_base=np.array([args.rope_base]), eps=np.array([args.norm_eps]))
    print("forward_llamagen.npz:", [l.shape for l in logits])


@torch.no_grad()
def mint_chameleon():
    from oracle.mint_golden import apply_shims
    Cache = apply_shims()
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(REF / "lumina_mgpt"))
    from model.chameleon.configuration_chameleon import ChameleonConfig
    ChameleonConfig.rope_scaling = None   # HF 5.5 auto-fills a dict the vendored code cannot parse
    from model.chameleon.modeling_chameleon import ChameleonForConditionalGeneration
    from oracle.mint_golden import load_reference_scheduler
    renew_backbone = load_reference_scheduler()[0].renew_backbone
    torch.manual_seed(5)
    cfg = ChameleonConfig(vocab_size=2048, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                          num_attention_heads=2, num_key_value_heads=2, max_position_embeddings=128,
                          rms_norm_eps=1e-5, mask_image_logits=False, attn_implementation="sdpa",
                          vocabulary_map={"<image>": 3, "IMGIMGA": 4, "IMGIMGB": 5},
                          vq_config={"embed_dim": 8, "num_embeddings": 16, "resolution": 32, "channel_multiplier": [1, 1],
                                     "base_channels": 32, "num_res_blocks": 1, "latent_channels": 8})
    m = ChameleonForConditionalGeneration(cfg).float().eval()
    from oracle import ref_forward as RF
    H, Dh = cfg.num_attention_heads, cfg.hidden_size // cfg.num_attention_heads
    rcfg = RF.StackConfig(cfg.num_hidden_layers, cfg.hidden_size, H, H, Dh, cfg.intermediate_size, cfg.vocab_size,
                          cfg.rms_norm_eps, qk_norm=True)
    w = RF.random_weights(rcfg, seed=CHAMELEON_SEED, std=0.05)
    m.model.embed_tokens.weight.copy_(w["embed"]); m.model.norm.weight.copy_(w["final_norm"])
    m.lm_head.weight.copy_(w["lm_head"])
    for L, wl in zip(m.model.layers, w["layers"]):
        a_, f_ = L.self_attn, L.mlp
        L.input_layernorm.weight.copy_(wl["attn_norm"]); L.post_attention_layernorm.weight.copy_(wl["ffn_norm"])
        hd = H * Dh
        a_.q_proj.weight.copy_(wl["wqkv"][:hd]); a_.k_proj.weight.copy_(wl["wqkv"][hd:2 * hd])
        a_.v_proj.weight.copy_(wl["wqkv"][2 * hd:]); a_.o_proj.weight.copy_(wl["wo"])
        f_.gate_proj.weight.copy_(wl["w_gate_up"][:cfg.intermediate_size])
        f_.up_proj.weight.copy_(wl["w_gate_up"][cfg.intermediate_size:]); f_.down_proj.weight.copy_(wl["w_down"])
        # the vendored ChameleonLayerNorm keeps [model_parallel_size, Dh] and repeat-interleaves over heads (:206-219):
        # size it to one row per head so that every head gets its own gamma/beta
        for mod, kw, kb in ((a_.q_norm, "q_norm_w", "q_norm_b"), (a_.k_norm, "k_norm_w", "k_norm_b")):
            assert mod.weight.shape[-1] == Dh
            if mod.weight.shape[0] == H:
                mod.weight.copy_(wl[kw]); mod.bias.copy_(wl[kb])
            else:   # fewer rows than heads: broadcast the first rows, and record what the heads really see
                r = mod.weight.shape[0] if mod.weight.dim() == 2 else 1
                mod.weight.copy_(wl[kw][:r].reshape(mod.weight.shape)); mod.bias.copy_(wl[kb][:r].reshape(mod.bias.shape))
    m.model.__class__ = renew_backbone(m.model.__class__)
    rows = 2
    g = torch.Generator().manual_seed(7)
    P = 9
    prompt = torch.randint(0, cfg.vocab_size, (1, P), generator=g).repeat(rows, 1)
    cache = Cache()
    # CFG-uncond row: prompt keys [0, P-1) hidden (jacobi_iteration_lumina_mgpt.py:755-758)
    mask2d = torch.ones(rows, P, dtype=torch.long)
    mask2d[1, : P - 1] = 0
    pos = (mask2d.cumsum(-1) - 1).clamp(min=0)
    calls, logits, ids_all = [], [], []
    out = m(input_ids=prompt, attention_mask=mask2d, position_ids=pos, past_key_values=cache, use_cache=True,
            cache_position=torch.arange(P))
    calls.append((0, P))
    ids_all.append(prompt.numpy())
    logits.append(out.logits.float().numpy())
    kv_len = P
    for W, rollback in ((5, 0), (7, 3), (1, 0)):
        if rollback:   # drop the last `rollback` cached keys, like delete_false_key_value (:47-54)
            kv_len -= rollback
            cache.key_cache = [k[:, :, :kv_len] for k in cache.key_cache]
            cache.value_cache = [v[:, :, :kv_len] for v in cache.value_cache]
        ids = torch.randint(0, cfg.vocab_size, (1, W), generator=g).repeat(rows, 1)
        T = kv_len + W
        m3 = torch.ones(rows, W, T, dtype=torch.long)      # 3-D window mask (:798-863)
        m3[1, :, : P - 1] = 0
        m3[:, :, kv_len:] = torch.tril(torch.ones(W, W, dtype=torch.long))
        cp = torch.arange(kv_len, T)
        pid = torch.stack([cp, cp - (P - 1)])
        out = m(input_ids=ids, attention_mask=m3, position_ids=pid, past_key_values=cache, use_cache=True,
                cache_position=cp)
        calls.append((kv_len, W))
        ids_all.append(ids.numpy())
        logits.append(out.logits.float().numpy())
        kv_len = T
    qk_rows = int(m.model.layers[0].self_attn.q_norm.weight.shape[0]) if m.model.layers[0].self_attn.q_norm.weight.dim() == 2 else 1
    np.savez_compressed(OUT / "forward_chameleon.npz", calls=np.array(calls), P=np.array([P]), seed=np.array([CHAMELEON_SEED]),
                        qk_norm_rows=np.array([qk_rows]),
                        **{f"ids{i}": a for i, a in enumerate(ids_all)}, **{f"logits{i}": a for i, a in enumerate(logits)},
                        cfg=np.array([cfg.num_hidden_layers, cfg.hidden_size, H, H, Dh, cfg.intermediate_size,
                                      cfg.vocab_size]),
                        theta=np.array([cfg.rope_theta if hasattr(cfg, "rope_theta") else 10000.0]),
                        eps=np.array([cfg.rms_norm_eps]))
    print("forward_chameleon.npz:", [l.shape for l in logits])


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    which = sys.argv[1:] or ["llamagen", "chameleon"]
    if "llamagen" in which:
        mint_llamagen()
    if "chameleon" in which:
        mint_chameleon()
