"""Helper of tests/test_gpu_baseline_sizes.py::test_flexar_inference_solver_flow_on_gpu (run as a subprocess so that the
HF-5.5 name shims the reference needs stay out of the pytest process).

Drives the reference's Lumina entry point the way test_lumina_mgpt.py does (:101-138):

    solver = FlexARInferenceSolver(...)                         lumina_mgpt/inference_solver.py:273-296
    solver = renew_pipeline_sampler(solver, **jacobi_params)    THIS repository's scheduler.jacobi_iteration_lumina_mgpt
    solver.generate(images=[], qas=[[q, None]], max_gen_len, temperature,
                    logits_processor=solver.create_logits_processor(cfg, image_top_k))     :298-354

FlexARInferenceSolver is the reference's OWN class when an install of the reference is present (baseline/_ref, see
scripts/install_reference.py) — constructed without its checkpoint-downloading __init__ — and otherwise a restatement
of the same generate() body.  The tokenizer / VQ item processor is a stub (no checkpoints offline); the model is a tiny
random-init HF ChameleonForConditionalGeneration on the GPU, so generate() goes HF GenerationMixin.generate -> the
renewed _sample -> SJD engine kernels.  Prints one JSON line.
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "baseline" / "_ref"
sys.dont_write_bytecode = True
sys.path.insert(0, str(ROOT))

GRID = 8
PROMPT = [0, 9000, 9001, 9002, 9003, 8197, 8804 + GRID // 2, 8804 + GRID // 2]


class StubTokenizer:
    def decode(self, ids):
        return " ".join(str(int(i)) for i in ids)


class StubItemProcessor:
    """What FlexARItemProcessor offers to generate()/decode_ids()/create_logits_processor (data/item_processor.py)."""
    image_start_token, image_end_token, new_line_token = "<racm3:break>", "<eoss>", "<reserved08796>"
    _ids = {"<racm3:break>": 8197, "<eoss>": 8196, "<reserved08796>": 8803, "<|image|>": 8711}
    tokenizer = StubTokenizer()

    def token2id(self, name):
        return self._ids[name]

    def process_item(self, item, **kw):
        assert item["conversations"][0]["from"] == "human"
        return list(PROMPT)

    def decode_image(self, tokens):
        return list(tokens)        # the VQ decoder is outside the SJD path: hand the latent ids back


def restated_solver_class():
    from transformers import GenerationConfig

    class FlexARInferenceSolver:   # restatement of lumina_mgpt/inference_solver.py:298-354 (+ decode_ids :356-398)
        @torch.no_grad()
        def generate(self, images, qas, max_gen_len, temperature, logits_processor=None, streamer=None):
            conversations = []
            for q, a in qas:
                conversations += [{"from": "human", "value": q}, {"from": "gpt", "value": a}]
            _prompt = self.item_processor.process_item({"image": images, "conversations": conversations})
            prompt = []
            for value in _prompt:
                prompt += [value] if isinstance(value, int) else value["input_ids"]
            prompt_len = len(prompt)
            prompt = torch.tensor(prompt, dtype=torch.int64, device=self.model.device).unsqueeze(0)
            generation_config = GenerationConfig(max_new_tokens=max_gen_len, max_length=self.model.config.max_position_embeddings,
                                                 temperature=temperature, top_k=None, do_sample=True, eos_token_id=[8710])
            if logits_processor is None:
                logits_processor = self.create_logits_processor()
            with torch.autocast("cuda", dtype=self.dtype):
                result = self.model.generate(prompt, generation_config, logits_processor=logits_processor,
                                             streamer=streamer)[0][prompt_len:].tolist()
                if len(result) > 0 and result[-1] == 8710:
                    result = result[:-1]
            return self.decode_ids(result)

        def decode_ids(self, tokens):
            ip = self.item_processor
            images, text, i = [], [], 0
            while i < len(tokens):
                if tokens[i] == ip.token2id(ip.image_start_token) and ip.token2id(ip.image_end_token) in tokens[i + 1:]:
                    j = tokens.index(ip.token2id(ip.image_end_token), i + 1)
                    images.append(ip.decode_image(tokens[i + 1:j]))
                    text.append(ip.token2id("<|image|>"))
                    i = j + 1
                else:
                    text.append(tokens[i])
                    i += 1
            return ip.tokenizer.decode(text), images

    return FlexARInferenceSolver


def main():
    from transformers import ChameleonConfig, ChameleonForConditionalGeneration
    dev = torch.device("cuda:0")
    cls_from = "restated"
    Solver = None
    if (REF / "lumina_mgpt" / "inference_solver.py").exists():
        try:
            sys.path[1:1] = [str(REF), str(REF / "lumina_mgpt")]
            from oracle.mint_golden import apply_shims
            apply_shims()
            from lumina_mgpt.inference_solver import FlexARInferenceSolver as Solver   # the reference's own class
            cls_from = "reference"
        except Exception as e:   # pragma: no cover
            print("reference class not importable here:", repr(e)[:200], file=sys.stderr)
            Solver = None
    if Solver is None:
        Solver = restated_solver_class()
    torch.manual_seed(3)
    cfg = ChameleonConfig(vocab_size=9216, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                          num_attention_heads=2, num_key_value_heads=2, max_position_embeddings=512, rms_norm_eps=1e-5,
                          vocabulary_map={"<image>": 3, "IMGIMGA": 4, "IMGIMGB": 5}, eos_token_id=8710, bos_token_id=0,
                          pad_token_id=1,
                          vq_config={"embed_dim": 8, "num_embeddings": 16, "resolution": 32, "channel_multiplier": [1, 1],
                                     "base_channels": 32, "num_res_blocks": 1, "latent_channels": 8})
    model = ChameleonForConditionalGeneration(cfg)
    with torch.no_grad():
        for n_, p_ in model.named_parameters():
            if "vqmodel" not in n_ and p_.dim() >= 2 and "norm" not in n_:
                p_.normal_(0.0, 0.08)
    solver = Solver.__new__(Solver)          # __init__ downloads checkpoints (inference_solver.py:282-289)
    solver.dtype, solver.device = torch.bfloat16, dev
    solver.model = model.to(dev, torch.bfloat16).eval()
    solver.item_processor = StubItemProcessor()

    # ---- test_lumina_mgpt.py:101-114, with THIS repository's scheduler module ----
    from scheduler.jacobi_iteration_lumina_mgpt import renew_pipeline_sampler
    assert str(ROOT / "scheduler") in sys.modules["scheduler.jacobi_iteration_lumina_mgpt"].__file__
    W = 8
    solver = renew_pipeline_sampler(solver, jacobi_loop_interval_l=3, jacobi_loop_interval_r=GRID * GRID + GRID - 10,
                                    max_num_new_tokens=W, guidance_scale=3.0, seed=5, multi_token_init_scheme="random",
                                    do_cfg=True, image_top_k=2000, text_top_k=10,
                                    prefix_token_sampler_scheme="speculative_jacobi")
    solver.model.seed = 5
    n_img = GRID * (GRID + 1)
    text, images = solver.generate(images=[], qas=[["Generate an image of 128x128 according to the following prompt:\nx", None]],
                                   max_gen_len=n_img + 1, temperature=1.0,
                                   logits_processor=solver.create_logits_processor(cfg=3.0, image_top_k=2000))
    st = solver.model.sjd_stats
    out = {"class_from": cls_from, "solver_class": type(solver).__mro__[1].__module__, "text": text, "n_images": len(images),
           "nfe": st.nfe, "new_tokens": st.new_tokens,
           "generated": None}
    print("RESULT" + json.dumps(out))
    # second call: the packed weights and the engine are reused, a parameter edit re-packs
    fp0 = id(solver.model._sjd_stack_cache)
    solver.generate(images=[], qas=[["again", None]], max_gen_len=20, temperature=1.0,
                    logits_processor=solver.create_logits_processor(cfg=3.0, image_top_k=2000))
    same = id(solver.model._sjd_stack_cache) == fp0
    with torch.no_grad():
        solver.model.lm_head.weight.mul_(1.0)
    solver.generate(images=[], qas=[["again", None]], max_gen_len=20, temperature=1.0,
                    logits_processor=solver.create_logits_processor(cfg=3.0, image_top_k=2000))
    print("CACHE" + json.dumps({"reused": same, "repacked_after_edit": id(solver.model._sjd_stack_cache) != fp0}))


if __name__ == "__main__":
    main()
