"""Multi-GPU = independent replicas, prompts sharded round-robin — exactly what the reference does with one
OS process per GPU (dataset_tools/multi_gpu_infer_with_prompt.py:146-172, multi_gpu_dataframe_split.py:31-63).
The SJD math has no collective; the only exchange is an all-gather of {new tokens, NFE, done} counters
(NCCL over NVLink on the box, gloo in the CPU tests) so rank 0 can report whole-job throughput.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_rank_world() -> tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend: str | None = None) -> tuple[int, int, int]:
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_prompts(n_prompts: int, rank: int, world: int) -> list[int]:
    """Prompt i runs on GPU i mod world."""
    return list(range(rank, n_prompts, world))


def gather_counters(new_tokens: int, nfe: int, done: int, device="cpu") -> torch.Tensor:
    """all-gather of one int64 triple per rank -> [world, 3] (identity when not distributed)."""
    t = torch.tensor([new_tokens, nfe, done], dtype=torch.int64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t[None]
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out)


def max_over_ranks(value: float, device="cpu") -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
