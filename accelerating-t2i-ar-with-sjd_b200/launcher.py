"""Multi-GPU prompt runner: the reference's dataset_tools/multi_gpu_infer_with_prompt.py (PromptWrapper.run :45-66,
run_caption_gen :69-132, one OS process per GPU :146-172) over torch.distributed.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_prompts.py \\
        --model synthetic/lumina-mgpt-7b-768 --prompts prompts.txt --output-dir workdir

Each rank owns one GPU and one full replica (the SJD path has no collective): it loads the model through
model_wrappers.model_loader.load_pretrained_model / get_forward_func, takes prompts rank, rank + world, ... (the
reference slices its dataframe the same way, multi_gpu_dataframe_split.py:31-63), skips prompts whose output file
exists (resume, :56-57), saves a tensor result as <idx>.pt and an image as <idx>.png (:58-64), and at the end the
per-rank counters {prompts done, new tokens, forwards, seconds} are all-gathered (NCCL on the box, gloo on CPU) so that
rank 0 can report whole-job throughput."""
from __future__ import annotations

import os
import time

import torch

from . import replicas


class PromptWrapper:
    """multi_gpu_infer_with_prompt.py:19-66"""

    def __init__(self, prompts, gpu_id=0, node_id=0, model_name="", output_dir="./workdir", seed=None):
        self.prompts, self.gpu_id, self.node_id, self.seed = list(prompts), gpu_id, node_id, seed
        self.model_name = model_name.split("/")[-1]
        self.output_dir = output_dir
        os.makedirs(output_dir, exist_ok=True)
        self.done = self.skipped = 0

    def run(self, sample_fn):
        for prompt_idx, prompt in self.prompts:
            path = os.path.join(self.output_dir, f"{prompt_idx}.png")
            if os.path.exists(path) or os.path.exists(path.replace(".png", ".pt")):
                self.skipped += 1
                continue
            result = sample_fn(prompt)
            if isinstance(result, torch.Tensor):
                torch.save(result, path.replace(".png", ".pt"))
            elif hasattr(result, "save"):
                result.save(path)
            else:
                raise ValueError(f"Invalid image type: {type(result)}")
            self.done += 1


def run_prompts(model_name, prompts, output_dir="./workdir", seed=None, loader=None, backend=None, **kwargs):
    """One rank of the job.  `prompts`: list of strings (prompt i keeps index i across ranks).  Returns the gathered
    counters [world, 4] = {prompts done, new tokens, forwards, milliseconds} (on every rank)."""
    rank, world, local = replicas.init_process_group(backend)
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        torch.cuda.set_device(local)
    device = torch.device("cuda", local) if use_cuda else torch.device("cpu")
    if loader is None:
        from model_wrappers.model_loader import get_forward_func, load_pretrained_model
        model = load_pretrained_model(model_name, device=device, seed=seed, **kwargs)
        forward = get_forward_func(model_name, model, **kwargs)
    else:
        model, forward = loader(model_name, device=device, seed=seed, **kwargs)
    mine = [(i, prompts[i]) for i in replicas.shard_prompts(len(prompts), rank, world)]
    tok = nfe = 0

    def counted(prompt):
        nonlocal tok, nfe
        out = forward(prompt)
        st = getattr(getattr(model, "engine", None), "stats", None)
        if st is not None:
            tok, nfe = tok + st.new_tokens, nfe + st.nfe
        return out

    wrap = PromptWrapper(mine, gpu_id=local, node_id=int(os.environ.get("GROUP_RANK", "0")), model_name=model_name,
                         output_dir=output_dir, seed=seed)
    replicas.barrier()
    t0 = time.perf_counter()
    with torch.no_grad():
        wrap.run(counted)
    if use_cuda:
        torch.cuda.synchronize()
    ms = int((time.perf_counter() - t0) * 1e3)
    t = torch.tensor([wrap.done, tok, nfe, ms], dtype=torch.int64, device=device if use_cuda else "cpu")
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and world > 1:
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        res = torch.stack(out).cpu()
    else:
        res = t[None].cpu()
    if rank == 0:
        dt = max(int(res[:, 3].max()), 1) / 1e3
        print(f"[launcher] {int(res[:, 0].sum())} prompts on {world} replica(s): {int(res[:, 1].sum())} tokens, "
              f"{int(res[:, 2].sum())} forwards, {dt:.2f} s (max over ranks) -> {int(res[:, 1].sum()) / dt:.1f} tokens/s")
    return res
