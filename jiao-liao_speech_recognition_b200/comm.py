"""Data-parallel communicator for the fine-tune step: one process per GPU, one in-place sum of the
flat fp32 adapter + lm_head gradient bucket per step through ``jl_comm_allreduce`` (NCCL over NVLink 5 / NVSwitch,
enqueued on the caller's stream, CUDA-graph capturable).  It stands where DDP's reducer + ProcessGroupNCCL stand in the
reference's stack (/root/reference/requirements.txt:1,75).  ``torch.distributed`` is only the out-of-band channel that
ships rank 0's 128-byte NCCL id to the other ranks."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib as L


class JLComm:
    """Owns one ``jl_comm`` handle.  ``JLComm.from_torch_distributed()`` builds it inside an initialised process group
    (any backend: gloo is enough, the id travels as a byte tensor); ``JLComm(id_bytes, rank, world)`` takes an id obtained
    elsewhere (``JLComm.unique_id()`` on rank 0)."""

    def __init__(self, id_bytes: bytes, rank: int, world: int):
        if len(id_bytes) != L.COMM_ID_BYTES:
            raise ValueError(f"NCCL id must be {L.COMM_ID_BYTES} bytes, got {len(id_bytes)}")
        lib = L.load()
        handle = C.c_void_p()
        buf = C.create_string_buffer(bytes(id_bytes), L.COMM_ID_BYTES)
        L.check(lib.jl_comm_init(C.cast(buf, C.c_void_p), rank, world, C.byref(handle)))
        self._h: Optional[C.c_void_p] = handle
        self.rank, self.world = rank, world

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(L.COMM_ID_BYTES)
        L.check(L.load().jl_comm_unique_id(C.cast(buf, C.c_void_p)))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, group=None) -> "JLComm":
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("JLComm.from_torch_distributed needs an initialised torch.distributed process group")
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return cls(box[0], rank, world)

    def allreduce_(self, buf: torch.Tensor) -> torch.Tensor:
        """In-place sum over ranks of a contiguous fp32 CUDA tensor, on the current stream."""
        if buf.dtype != torch.float32 or not buf.is_cuda or not buf.is_contiguous():
            raise ValueError("allreduce_ needs a contiguous fp32 CUDA tensor")
        if self._h is None:
            raise RuntimeError("communicator already destroyed")
        L.check(L.load().jl_comm_allreduce(self._h, buf.data_ptr(), buf.numel(), torch.cuda.current_stream().cuda_stream))
        return buf

    def destroy(self) -> None:
        if self._h is not None:
            h, self._h = self._h, None
            L.check(L.load().jl_comm_destroy(h))

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
