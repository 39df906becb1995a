"""Data parallelism for the GenRL step: one process per GPU, batch sharded by B, gradients averaged over NCCL/NVLink.

Sequences are independent and every loss is a mean over per-row terms, so with equal shards the averaged gradient is
the global-batch gradient.  Non-gradient state that needs the global batch: the RewardEMA quantiles
(dreamer_utils.py:1022-1026) -> all-gather of the lambda-return targets.
Works with any torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests of the host logic).
"""
import torch
import torch.distributed as dist


class DataParallel:
    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.nccl = dist.get_backend(group) == "nccl"
        import os
        self.bf16_transport = os.environ.get("GENRL_DP_BF16", "0") == "1"
        self.n_allreduce = 0
        self.bytes_allreduced = 0

    def allreduce(self, flats):
        """average flat fp32 gradient buffers in place: one collective per optimizer-group buffer, issued on the CURRENT stream (the
        agent calls this on the stream whose backward produced the gradients, so that a finished group's collective runs beside the
        other stream's remaining backward)"""
        if self.world == 1:
            return
        for g in flats:
            if self.nccl and g.is_cuda and self.bf16_transport and g.dtype == torch.float32:
                # half the bytes on the wire: round to bf16 (nearest even, unbiased), average, widen again (GENRL_DP_BF16=1)
                h = g.to(torch.bfloat16)
                dist.all_reduce(h, op=dist.ReduceOp.AVG, group=self.group)
                g.copy_(h)
            elif self.nccl and g.is_cuda:
                dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
                g.div_(self.world)
            self.n_allreduce += 1
            self.bytes_allreduced += g.numel() * g.element_size()

    def gather(self, flat):
        """concatenate a 1-D tensor from every rank (equal lengths)"""
        if self.world == 1:
            return flat
        if not self.nccl and flat.is_cuda:          # gloo has no CUDA all-gather: stage through the host (tests on a single GPU)
            parts = [torch.empty(flat.numel(), dtype=flat.dtype) for _ in range(self.world)]
            dist.all_gather(parts, flat.detach().cpu().contiguous(), group=self.group)
            return torch.cat(parts).to(flat.device)
        out = torch.empty(self.world * flat.numel(), device=flat.device, dtype=flat.dtype)
        dist.all_gather_into_tensor(out, flat.contiguous(), group=self.group)
        return out

    def shard(self, data):
        """rows [r*B/world, (r+1)*B/world) of every tensor of a global batch dict"""
        out = {}
        for k, v in data.items():
            B = v.shape[0]
            assert B % self.world == 0, f"global batch {B} does not divide over {self.world} ranks"
            per = B // self.world
            out[k] = v[self.rank * per:(self.rank + 1) * per]
        return out

    def attach(self, agent):
        """route the agent's optimizer steps and RewardEMA through this group"""
        agent.dp = self
        agent._store_sig = None          # optimizers pick up the hook at the next ensure_store()
        for attr in ("_acting_behavior", "_imag_behavior"):
            beh = getattr(agent, attr, None)
            if beh is not None:
                beh.dp_gather = self.gather
        return agent
