"""Multi-GPU driver of the hot path: one process per GPU, each owning a shard of the quartet rank space.

SURVEY.md §8e / DESIGN.md §5.  Quartets are independent given the gene trees' distance matrices, so the
rank space is split by the outer (largest) taxon id s3 into G contiguous ranges of equal size
(``qs_shard_bounds``); every rank receives all flattened gene trees, builds all distance matrices itself and
counts only its own quartets -- no count traffic.  The only exchange is

* ``all_reduce(MIN)`` of the per-edge LQ-IC partials (edge_count doubles), and
* ``all_reduce(SUM)`` of the per-inner-node-pair topology sums (3 x n_pairs int64) -- QP-IC/EQP-IC are
  non-linear in those sums (``log_score``, src/QuartetScoreComputer.hpp:135-159, :472), so the SUMS are
  reduced and every rank finalises locally,

over NCCL (NVLink/NVSwitch) on a GPU box, or gloo in the CPU tests of this logic.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _ffi
from .computer import Context


def shard_bounds(n_taxa: int, shard_index: int, shard_count: int) -> Tuple[int, int, int, int]:
    """(s3_begin, s3_end, rank_begin, rank_end) of a shard -- the library's own rule (include/qscuda.h)."""
    lib = _ffi.load()
    b, e = C.c_int(), C.c_int()
    rb, re_ = C.c_uint64(), C.c_uint64()
    rc = lib.qs_shard_bounds(n_taxa, shard_index, shard_count, C.byref(b), C.byref(e), C.byref(rb), C.byref(re_))
    if rc != 0:
        raise _ffi.QSError(rc, "qs_shard_bounds: " + lib.qs_strerror(rc).decode())
    return b.value, e.value, rb.value, re_.value


def allreduce_partials(lqic_partial: np.ndarray, pair_sums: np.ndarray, group=None) -> Tuple[np.ndarray, np.ndarray]:
    """MIN-reduce the LQ-IC partials and SUM-reduce the pair sums over the default (or given) process group.

    Works with the nccl backend (tensors staged through the current CUDA device) and with gloo (CPU tensors).
    uint64 sums travel as int64 (two's complement; the sums of any realistic run are far below 2^63)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return lqic_partial, pair_sums
    on_gpu = dist.get_backend(group) == "nccl"
    t_lq = torch.from_numpy(np.ascontiguousarray(lqic_partial, np.float64))
    t_s = torch.from_numpy(np.ascontiguousarray(pair_sums, np.uint64).view(np.int64))
    if on_gpu:
        t_lq, t_s = t_lq.cuda(non_blocking=True), t_s.cuda(non_blocking=True)
    dist.all_reduce(t_lq, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(t_s, op=dist.ReduceOp.SUM, group=group)
    return t_lq.cpu().numpy(), t_s.cpu().numpy().view(np.uint64)


class _DeviceArray:
    """int64 view of `count` elements at a raw device address, for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def score_distributed(ctx: Context, count_scale: int = 1, exact_qp: bool = False, group=None):
    """Scores of the WHOLE quartet space from a sharded context.

    NCCL: the per-pair partials never leave the device -- local scan, all-reduce(SUM) of the topology sums,
    all-reduce(MIN) of the per-pair minimal scores, all-reduce(MIN) of the winners' count triples, all in place on the
    library's own device buffers and on the context's stream (which must be torch's current stream); then the per-edge
    reduction on the device and log_score of ~3 x edges values on the host.  Other backends (gloo in the CPU tests of the
    host logic): host partials as before."""
    if ctx.shard_count == 1:
        return ctx.score(count_scale, exact_qp)
    import torch
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_backend(group) == "nccl":
        ctx.score_scan(count_scale)
        p_sums, p_score, p_best, n_pairs = ctx.score_device_partials()
        dev = torch.device("cuda", ctx.device)
        t_sums = torch.as_tensor(_DeviceArray(p_sums, 3 * n_pairs), device=dev)
        t_score = torch.as_tensor(_DeviceArray(p_score, n_pairs), device=dev)
        t_best = torch.as_tensor(_DeviceArray(p_best, n_pairs), device=dev)
        dist.all_reduce(t_sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(t_score, op=dist.ReduceOp.MIN, group=group)
        ctx.score_select_winners()
        dist.all_reduce(t_best, op=dist.ReduceOp.MIN, group=group)
        return ctx.score_finish(exact_qp)
    lq, sums = ctx.score_partials(count_scale)
    lq, sums = allreduce_partials(lq, sums, group)
    return ctx.score_finalize(lq, sums, exact_qp)


def finalize_on_host(ref, n_taxa: int, lqic_reduced: np.ndarray, pair_sums_reduced: np.ndarray, exact_qp: bool = False,
                     cint_bytes: int = 2):
    """Turn reduced partials into (lqic, qpic, eqpic) on a rank that holds no GPU context (QS_DEVICE_NONE)."""
    with Context(n_taxa, cint_bytes, device=_ffi.QS_DEVICE_NONE) as host:
        host.set_reference(ref)
        return host.score_finalize(lqic_reduced, pair_sums_reduced, exact_qp)
