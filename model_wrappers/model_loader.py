"""Drop-in for the reference's model_wrappers/model_loader.py: `load_pretrained_model(model_name, **kw)` and
`get_forward_func(model_name, model, **kw)` — the two calls the reference's multi-GPU prompt runner makes per process
(dataset_tools/multi_gpu_infer_with_prompt.py:110-122) — over the sm_100a SJD engine.

reference (model_loader.py)                       here
----------------------------------------------------------------------------------------------------------------------
load_lumina_mgpt :25-60, forward func :362-387    same steps: FlexARInferenceSolver (the reference's tokenizer + VQ item
                                                  processor, its own code) -> THIS repository's renew_pipeline_sampler
load_anole :62-110, forward func :389-425         HF ChameleonForConditionalGeneration + ChameleonProcessor ->
                                                  scheduler.jacobi_iteration_anhole.renew_pipeline_sampler (this repository)
load_emu3 :112-192, forward func :427-496         AutoModelForCausalLM + Emu3Processor -> scheduler.jacobi_iteration_emu3.renew_solver
load_llamagen :194-345, forward func :498-562     GPT_models + VQ + T5 -> renew_llamagen / renew_sampler / LlamaGenSolver
get_forward_func / load_pretrained_model :347-360, :564-574     same dispatch on the model name

Differences, on purpose:
  * the LlamaGen forward function decodes with the Jacobi solver (LlamaGenSolver.generate).  The reference's calls the
    plain AR `generate` there (model_loader.py:544) although it has just built the solver — its LlamaGen numbers through
    eval_model.py are therefore NOT SJD numbers.  `use_jacobi=False` reproduces that.
  * `model_name = "synthetic/<family>"` (lumina-mgpt-7b-768, anole-7b-512, emu3-gen, llamagen-gpt-b) builds the
    family's decoder stack with random-init weights directly in the engine's layout: no checkpoint, tokenizer or VQ
    decoder (none are reachable offline), prompts are hashed to synthetic token ids and the forward function returns the
    generated token ids (a LongTensor — what the reference returns with not_decoded_imgs=True, :447-478).  bench.py and
    the multi-GPU launcher (sjd_b200.launcher) run on these.
Real checkpoints need the reference's tokenizer / VQ code on sys.path after this repository and network access, like
upstream.
"""
from __future__ import annotations

import hashlib

import torch

_SYNTH = "synthetic/"


# ------------------------------------------------------------------------------------------- synthetic families
def _synth_prompt_ids(prompt: str, n: int, lo: int, hi: int) -> list[int]:
    seed = int.from_bytes(hashlib.sha256(prompt.encode()).digest()[:8], "little")
    g = torch.Generator().manual_seed(seed % (2 ** 63))
    return torch.randint(lo, hi, (n,), generator=g).tolist()


class SyntheticSolver:
    """One family's engine + grammar behind `generate(prompt) -> LongTensor[new tokens]`."""

    def __init__(self, family, stack, engine_obj, make_prompt, max_new, eos, kv_lo_fn, name):
        self.family, self.stack, self.engine, self.name = family, stack, engine_obj, name
        self._make_prompt, self.max_new, self.eos, self._kv_lo = make_prompt, max_new, eos, kv_lo_fn
        self.device = stack.device

    @torch.no_grad()
    def generate(self, prompt: str, seed=None):
        ids = self._make_prompt(prompt)
        if seed is not None:
            self.engine.p.seed = seed
        out = self.engine.generate(ids, max_length=len(ids) + self.max_new, eos_token_ids=self.eos, kv_lo=self._kv_lo(ids))
        return torch.tensor(out[len(ids):], dtype=torch.long)


def _load_synthetic(model_name, device="cuda:0", seed=1, max_num_new_tokens=16, multi_token_init_scheme="random",
                    guidance_scale=3.0, image_top_k=2000, text_top_k=10, prefix_token_sampler_scheme="speculative_jacobi",
                    target_size=None, n_layers=None, weights_seed=0, **kwargs):
    import sjd_b200  # noqa: F401
    from sjd_b200 import engine as E, families, model
    fam = model_name[len(_SYNTH):].lower()
    dev = torch.device(device)
    if "lumina-mgpt" in fam or "anole" in fam:
        size = target_size or (512 if "anole" in fam or "512" in fam else 768)
        grid = size // 16
        shape = families.lumina_7b()
        theta, n_text = 10000.0, 64
        if "anole" in fam:   # HF Chameleon: image_seq_length = grid^2 ids in [4, 8196), no end-of-line tokens (Anole grammar)
            S = grid * grid
            max_new = S + 2
            grammar = E.AnoleGrammarState(8197, 8196, 2, 4, 8196, S, max_length=n_text + S + 2, begin_index=n_text,
                                          top_k=image_top_k)
            make_prompt = lambda p: _synth_prompt_ids(p, n_text, 9000, 60000)                      # noqa: E731
            interval_r, eos = S + 4, [2]
        else:
            max_new = grid * (grid + 1) + 2
            grammar = E.LuminaGrammarState(image_top_k=image_top_k, text_top_k=text_top_k)
            make_prompt = lambda p: _synth_prompt_ids(p, n_text, 9000, 60000) + [8197, 8804 + grid // 2, 8804 + grid // 2]  # noqa: E731
            interval_r, eos = grid * grid + grid - 10, [8710]
        img_vocab = torch.arange(4, 8196)
        kv_lo_fn = lambda ids: [0, len(ids) - 1]                                                   # noqa: E731
        max_len = n_text + 3 + max_new + max_num_new_tokens + 8
    elif "emu3" in fam:
        size = target_size or 720
        grid = size // 8
        shape = families.emu3_gen()
        theta, n_text = 1e6, 32
        vis_lo, vis_hi = 151854, 151854 + 32768
        img_tok, eol, eof, eoi, eos_t, pad = 151851, 151846, 151847, 151853, 151850, 151643
        grammar = E.Emu3GrammarState(grid, grid, img_tok, eol, eof, eoi, eos_t, pad, vis_lo, vis_hi, top_k=image_top_k)
        make_prompt = lambda p: _synth_prompt_ids(p, n_text, 1000, 150000) + [img_tok]              # noqa: E731
        max_new = grid * (grid + 1) + 3
        interval_r, eos = max_new + 8, [eos_t]
        img_vocab = torch.arange(vis_lo, vis_hi)
        kv_lo_fn = lambda ids: [0, 0]                                                              # noqa: E731
        max_len = n_text + 1 + max_new + max_num_new_tokens + 8
    elif "llamagen" in fam:
        name = "GPT-" + (fam.split("gpt-")[1].upper() if "gpt-" in fam else "B")
        size = target_size or 256
        grid = size // 16
        shape = families.llamagen(name)
        theta, n_text = None, 1
        grammar = E.PlainTopKState(top_k=image_top_k)
        make_prompt = lambda p: _synth_prompt_ids(p, 1 + 1, 0, shape.vocab)       # "class token" + first image token  # noqa: E731
        max_new = grid * grid - 1
        interval_r, eos = grid * grid - max_num_new_tokens - 2, []
        img_vocab = torch.arange(shape.vocab)
        kv_lo_fn = lambda ids: [0, 0]                                                              # noqa: E731
        max_len = grid * grid + 2
    else:
        raise NotImplementedError(f"unknown synthetic family {fam!r}")
    if n_layers:
        shape.n_layers = int(n_layers)
    max_len = int(-(-max_len // 64) * 64)
    w = families.random_weights(shape, seed=weights_seed, device=dev)
    if theta is None:
        cos, sin = families.rope_llamagen_2d(grid, shape.head_dim, 10000.0, 1)
        max_len = max(max_len, cos.shape[0])
    else:
        cos, sin = families.rope_rotate_half(shape.head_dim, max_len, theta, True)
    stack = model.DeviceStack(shape, w, cos, sin, rows=2, max_len=max_len, device=dev)
    del w
    params = E.SJDParams(jacobi_loop_interval_l=1, jacobi_loop_interval_r=interval_r, max_num_new_tokens=max_num_new_tokens,
                         guidance_scale=guidance_scale, seed=seed, multi_token_init_scheme=multi_token_init_scheme,
                         do_cfg=True, prefix_token_sampler_scheme=prefix_token_sampler_scheme)
    eng = E.SJDEngine(stack, params, grammar, img_vocab)
    return SyntheticSolver(fam, stack, eng, make_prompt, max_new, eos, kv_lo_fn, model_name)


# ------------------------------------------------------------------------------------------- real checkpoints
def load_lumina_mgpt(cache_dir="./ckpts", model_name="Alpha-VLLM/Lumina-mGPT-7B-768", target_size=768, seed=1,
                     max_num_new_tokens=16, multi_token_init_scheme="random", guidance_scale=7.0, device="cpu", **kwargs):
    """model_loader.py:25-60"""
    from lumina_mgpt.inference_solver import FlexARInferenceSolver          # the reference's own tokenizer / VQ code
    from scheduler.jacobi_iteration_lumina_mgpt import renew_pipeline_sampler
    solver = FlexARInferenceSolver(model_path=model_name, precision="bf16", target_size=target_size, cache_dir=cache_dir,
                                   device=device)
    return renew_pipeline_sampler(solver, jacobi_loop_interval_l=1,
                                  jacobi_loop_interval_r=(target_size // 16) ** 2 + target_size // 16 - 10,
                                  max_num_new_tokens=max_num_new_tokens, guidance_scale=guidance_scale, seed=seed,
                                  multi_token_init_scheme=multi_token_init_scheme, do_cfg=True, **kwargs)


def load_anole(cache_dir="./ckpts", model_name="leloy/Anole-7b-v0.1-hf", seed=1, max_num_new_tokens=16,
               multi_token_init_scheme="random", guidance_scale=3.0, device="cpu", image_top_k=2000, **kwargs):
    """model_loader.py:62-110"""
    from transformers import ChameleonForConditionalGeneration, ChameleonProcessor
    from scheduler.jacobi_iteration_anhole import renew_pipeline_sampler
    processor = ChameleonProcessor.from_pretrained(model_name, cache_dir=cache_dir)
    model = ChameleonForConditionalGeneration.from_pretrained(model_name, device_map=device, torch_dtype=torch.bfloat16,
                                                              cache_dir=cache_dir)
    S = int(getattr(processor, "image_seq_length", 1024))
    model = renew_pipeline_sampler(model, processor, jacobi_loop_interval_l=1, jacobi_loop_interval_r=S + 4,
                                   max_num_new_tokens=max_num_new_tokens, guidance_scale=guidance_scale, seed=seed,
                                   multi_token_init_scheme=multi_token_init_scheme, do_cfg=True, image_top_k=image_top_k,
                                   **kwargs)
    return {"model": model, "processor": processor}


def load_emu3(cache_dir="./ckpts", model_name="BAAI/Emu3-Gen", vq_hub="BAAI/Emu3-VisionTokenizer", seed=1,
              max_num_new_tokens=16, multi_token_init_scheme="random", guidance_scale=3.0, device="cpu", image_top_k=2048,
              target_size=720, **kwargs):
    """model_loader.py:112-192"""
    from transformers import AutoImageProcessor, AutoModel, AutoModelForCausalLM, AutoTokenizer
    from transformers.generation.configuration_utils import GenerationConfig
    from emu3.mllm.processing_emu3 import Emu3Processor                      # the reference's own processor
    from scheduler.jacobi_iteration_emu3 import renew_solver
    model = AutoModelForCausalLM.from_pretrained(model_name, device_map=device, torch_dtype=torch.bfloat16,
                                                 attn_implementation="sdpa", trust_remote_code=True, cache_dir=cache_dir)
    tokenizer = AutoTokenizer.from_pretrained(model_name, trust_remote_code=True, padding_side="left", cache_dir=cache_dir)
    image_processor = AutoImageProcessor.from_pretrained(vq_hub, trust_remote_code=True, cache_dir=cache_dir)
    image_tokenizer = AutoModel.from_pretrained(vq_hub, device_map=device, trust_remote_code=True, cache_dir=cache_dir).eval()
    processor = Emu3Processor(image_processor, image_tokenizer, tokenizer)
    kw = dict(mode="G", ratio="1:1", image_area=target_size * target_size, return_tensors="pt", padding="longest")
    probe = processor(text="x", **kw)
    h, w = probe.image_size[0]
    gen_cfg = GenerationConfig(use_cache=True, eos_token_id=model.config.eos_token_id, pad_token_id=model.config.pad_token_id,
                               max_new_tokens=40960, do_sample=True, top_k=image_top_k)
    model, logits_processor = renew_solver(model, processor, h=h, w=w, jacobi_loop_interval_l=1,
                                           jacobi_loop_interval_r=(h + 1) * w + 16, max_num_new_tokens=max_num_new_tokens,
                                           guidance_scale=guidance_scale, seed=seed,
                                           multi_token_init_scheme=multi_token_init_scheme, do_cfg=True, **kwargs)
    return {"model": model, "processor": processor, "GENERATION_CONFIG": gen_cfg, "logits_processor": logits_processor,
            "processor_kwargs": kw}


def load_llamagen(cache_dir="./ckpts", model_name="llamagen-GPT-XL", gpt_ckpt=None, vq_ckpt=None, t5_path=None,
                  gpt_model="GPT-XL", gpt_type="t2i", image_size=512, downsample_size=16, cls_token_num=120,
                  codebook_size=16384, codebook_embed_dim=8, seed=1, max_num_new_tokens=16,
                  multi_token_init_scheme="repeat_horizon", guidance_scale=7.5, temperature=1.0, image_top_k=1000,
                  image_top_p=1.0, device="cpu", no_left_padding=False, precision=torch.bfloat16, **kwargs):
    """model_loader.py:194-345"""
    from llamagen.language.t5 import T5Embedder                              # the reference's own (needs ftfy, T5 weights)
    from llamagen.llamagen import GPT_models
    from llamagen.llamagen_solver import LlamaGenSolver, renew_llamagen
    from llamagen.tokenizer.tokenizer_image.vq_model import VQ_models        # the reference's own VQ decoder
    from scheduler.jacobi_iteration_lumina_mgpt import renew_sampler
    latent = image_size // downsample_size
    vq = VQ_models["VQ-16"](codebook_size=codebook_size, codebook_embed_dim=codebook_embed_dim).to(device).eval()
    vq.load_state_dict(torch.load(vq_ckpt, map_location="cpu")["model"])
    gpt = GPT_models[gpt_model](block_size=latent ** 2, cls_token_num=cls_token_num, model_type=gpt_type).to(device=device,
                                                                                                            dtype=precision)
    jd = dict(jacobi_loop_interval_l=1, jacobi_loop_interval_r=latent ** 2 - max_num_new_tokens - 2,
              max_num_new_tokens=max_num_new_tokens, guidance_scale=guidance_scale, seed=seed,
              multi_token_init_scheme=multi_token_init_scheme, do_cfg=True, image_top_k=image_top_k, text_top_k=10,
              prefix_token_sampler_scheme=kwargs.get("prefix_token_sampler_scheme", "speculative_jacobi"))
    gpt.__class__ = renew_llamagen(gpt.__class__)
    gpt._init_new_params(**jd)
    gpt.__class__ = renew_sampler(gpt.__class__)
    gpt._init_new_params(**jd)
    ck = torch.load(gpt_ckpt, map_location="cpu")
    gpt.load_state_dict(ck.get("model", ck.get("module", ck.get("state_dict", ck))), strict=False)
    gpt.eval()
    t5 = T5Embedder(device=device, local_cache=True, cache_dir=t5_path, dir_or_name="flan-t5-xl", torch_dtype=precision,
                    model_max_length=cls_token_num)
    solver = LlamaGenSolver(model=gpt, image_top_k=image_top_k, image_top_p=image_top_p)
    return {"model": solver, "gpt_model": gpt, "t5_model": t5, "vq_model": vq, "latent_size": latent,
            "vq_params": {"codebook_embed_dim": codebook_embed_dim}, "backbone_params": {"no_left_padding": no_left_padding},
            "guidance_scale": guidance_scale, "temperature": temperature, "image_top_k": image_top_k,
            "image_top_p": image_top_p}


def load_pretrained_model(model_name="Alpha-VLLM/Lumina-mGPT-7B-768", **kwargs):
    """model_loader.py:347-360 (+ the synthetic families)."""
    if model_name.startswith(_SYNTH):
        return _load_synthetic(model_name, **kwargs)
    name = model_name.lower()
    if "lumina-mgpt" in name:
        return load_lumina_mgpt(model_name=model_name, **kwargs)
    if "anole" in name:
        return load_anole(model_name=model_name, **kwargs)
    if "llamagen" in name:
        return load_llamagen(model_name=model_name, **kwargs)
    if "emu3" in name:
        return load_emu3(model_name=model_name, **kwargs)
    raise NotImplementedError


# ------------------------------------------------------------------------------------------- forward functions
def get_lumina_mgpt_forward_func(inference_solver, guidance_scale=7.0, image_top_k=2000, max_gen_len=8192, temperature=1.0,
                                 target_size=768, **kwargs):
    """model_loader.py:362-387"""
    def sample_fn(prompts):
        q = f"Generate an image of {target_size}x{target_size} according to the following prompt:\n" + prompts
        generated = inference_solver.generate(
            images=[], qas=[[q, None]], max_gen_len=max_gen_len, temperature=temperature,
            logits_processor=inference_solver.create_logits_processor(cfg=guidance_scale, image_top_k=image_top_k))
        return inference_solver.create_image_grid([generated[1][0]], 1, 1)
    return sample_fn


def get_anole_forward_func(inference_solver, **kwargs):
    """model_loader.py:389-425"""
    from PIL import Image
    processor, model = inference_solver["processor"], inference_solver["model"]

    def sample_fn(prompts):
        inputs = processor("Generate an image of " + prompts, padding=True, return_tensors="pt").to(model.device, dtype=model.dtype)
        ids = model.generate(**inputs, multimodal_generation_mode="image-only", max_new_tokens=1026, do_sample=True)
        resp = ids[:, inputs["input_ids"].shape[-1]:]
        px = processor.postprocess_pixel_values(model.decode_image_tokens(resp[:, 1:-1]))
        return Image.fromarray(px[0].permute(1, 2, 0).cpu().numpy().astype("uint8"))
    return sample_fn


def get_emu3_forward_func(inference_solver, not_decoded_imgs=False, **kwargs):
    """model_loader.py:427-496"""
    from PIL import Image
    processor, model = inference_solver["processor"], inference_solver["model"]
    gen_cfg, logits_processor = inference_solver["GENERATION_CONFIG"], inference_solver["logits_processor"]
    pkw = inference_solver.get("processor_kwargs", kwargs)
    POS = " masterpiece, film grained, best quality."
    NEG = ("lowres, bad anatomy, bad hands, text, error, missing fingers, extra digit, fewer digits, cropped, worst quality, "
           "low quality, normal quality, jpeg artifacts, signature, watermark, username, blurry.")

    def sample_fn(prompts):
        pos, neg = processor(text=prompts + POS, **pkw), processor(text=NEG, **pkw)
        dev = model.device
        pos_ids, neg_ids = torch.as_tensor(pos.input_ids).to(dev), torch.as_tensor(neg.input_ids).to(dev)
        mi = model.prepare_batch_cfg_model_inputs(pos_ids, neg_input_ids=neg_ids, attention_mask=None)
        out = model.generate(mi["pos_input_ids"], gen_cfg, logits_processor=logits_processor,
                             attention_mask=mi["attention_mask"], neg_input_ids=neg_ids)[0]
        if not_decoded_imgs:
            return out
        with torch.no_grad():
            imgs = [im for im in processor.decode(out) if isinstance(im, Image.Image)]
        return imgs[-1]
    return sample_fn


def get_llamagen_forward_func(inference_solver, use_jacobi=True, **kwargs):
    """model_loader.py:498-562.  use_jacobi=False: the reference's behaviour (plain AR `generate`, :544)."""
    from PIL import Image
    from llamagen.llamagen_solver import generate as llamagen_original_generate
    s = inference_solver
    solver, gpt, t5, vq, latent = s["model"], s["gpt_model"], s["t5_model"], s["vq_model"], s["latent_size"]
    cdim, no_left = s["vq_params"]["codebook_embed_dim"], s["backbone_params"]["no_left_padding"]

    def sample_fn(prompts):
        embs, masks = t5.get_text_embeddings([prompts])
        if not no_left:   # left-pad the caption (:522-533)
            new_masks = torch.flip(masks, dims=[-1])
            embs = torch.stack([torch.cat([e[int(m.sum()):], e[:int(m.sum())]]) for e, m in zip(embs, masks)])
            masks = new_masks
        c = embs * masks[:, :, None]
        kw = dict(cfg_scale=s["guidance_scale"], temperature=s["temperature"], top_k=s["image_top_k"], top_p=s["image_top_p"],
                  sample_logits=True)
        idx = (solver.generate(c, latent ** 2, masks, **kw) if use_jacobi
               else llamagen_original_generate(gpt, c, latent ** 2, masks, **kw))
        img = vq.decode_code(idx, [len(c), cdim, latent, latent]).clamp(min=-1, max=1)
        img = (img - img.min()) / (img.max() - img.min()) * 255
        return Image.fromarray(img[0].permute(1, 2, 0).cpu().numpy().astype("uint8"))
    return sample_fn


def get_synthetic_forward_func(solver, **kwargs):
    def sample_fn(prompts):
        return solver.generate(prompts)
    return sample_fn


def get_forward_func(model_name, model, **kwargs):
    """model_loader.py:564-574"""
    if model_name.startswith(_SYNTH):
        return get_synthetic_forward_func(model, **kwargs)
    name = model_name.lower()
    if "lumina-mgpt" in name:
        return get_lumina_mgpt_forward_func(model, **kwargs)
    if "anole" in name:
        return get_anole_forward_func(model, **kwargs)
    if "llamagen" in name:
        return get_llamagen_forward_func(model, **kwargs)
    if "emu3" in name:
        return get_emu3_forward_func(model, **kwargs)
    raise NotImplementedError
