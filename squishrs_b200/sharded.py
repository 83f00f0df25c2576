"""Dedup across the GPUs of one box: the digest index is sharded by digest prefix, one process per GPU.

This is the host-side plumbing of SURVEY.md section 8(e) / north_star (2): per batch each rank routes its
{digest, global chunk index} records to their owner ranks, an all-to-all (torch.distributed: NCCL over
NVLink on GPUs, gloo in the CPU tests) carries them, owners insert and answer with one verdict byte per
record, a second all-to-all brings the verdicts home.  The three device steps are C-ABI calls
(sq_route_digests_device / sq_dedup_insert_routed_device / sq_unroute_verdicts_device); `ops` abstracts
them so the exchange logic can be exercised with world_size 2 on CPU.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

REC_BYTES = 32  # {digest[16], gidx u64, pad u64}
INVALID = (1 << 64) - 1


def owner_of(digest: bytes, world: int) -> int:
    """owner rank of a digest: little-endian u64 of its first 8 bytes, modulo the world size"""
    return int.from_bytes(digest[:8], "little") % world


class DeviceOps:
    """The real thing: CUDA kernels behind the C ABI."""

    def __init__(self, ctx, stream_ptr):
        self.ctx, self.lib, self.sp = ctx, ctx.lib, stream_ptr

    def route(self, digests: torch.Tensor, gidx_base: int, n: int, world: int, cap: int, send: torch.Tensor, send_pos: torch.Tensor):
        self.ctx.check(self.lib.sq_route_digests_device(self.ctx.h, digests.data_ptr(), gidx_base, n, world, cap, send.data_ptr(),
                                                        send_pos.data_ptr(), self.sp))

    def insert(self, recv: torch.Tensor, count: int, verdict: torch.Tensor):
        self.ctx.check(self.lib.sq_dedup_insert_routed_device(self.ctx.h, recv.data_ptr(), count, verdict.data_ptr(), self.sp))

    def unroute(self, verdict_back: torch.Tensor, send_pos: torch.Tensor, n: int, is_new: torch.Tensor):
        self.ctx.check(self.lib.sq_unroute_verdicts_device(self.ctx.h, verdict_back.data_ptr(), send_pos.data_ptr(), n, is_new.data_ptr(), self.sp))


class ShardedDedup:
    """Per-rank driver of the routed dedup exchange.  Buffers are allocated once for batches of <= cap chunks."""

    def __init__(self, ops, world: int, cap: int, device):
        self.ops, self.world, self.cap = ops, world, cap
        self.send = torch.empty(world * cap * REC_BYTES, dtype=torch.uint8, device=device)
        self.recv = torch.empty_like(self.send)
        self.send_pos = torch.empty(cap, dtype=torch.int32, device=device)
        self.verdict = torch.empty(world * cap, dtype=torch.uint8, device=device)
        self.verdict_back = torch.empty_like(self.verdict)

    def exchange(self, digests: torch.Tensor, gidx_base: int, n: int, is_new: torch.Tensor, group=None):
        """digests: n*16 bytes of this rank's batch; is_new: n bytes out.  Collective: every rank must call."""
        assert n <= self.cap
        self.ops.route(digests, gidx_base, n, self.world, self.cap, self.send, self.send_pos)
        dist.all_to_all_single(self.recv, self.send, group=group)           # digests -> owners
        self.ops.insert(self.recv, self.world * self.cap, self.verdict)
        dist.all_to_all_single(self.verdict_back, self.verdict, group=group)  # verdicts -> senders
        self.ops.unroute(self.verdict_back, self.send_pos, n, is_new)
