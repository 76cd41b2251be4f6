"""NVLink peer-memory fabric: the data plane that replaces per-variable RecvTensor RPCs (SURVEY §5
"Distributed communication backend").

Every task allocates named device buffers; :meth:`Fabric.exchange` makes each buffer addressable
from every other task's kernels:

* **multi-process** (between-graph replication, one process per GPU under ``torchrun``): buffers are
  ``cudaMalloc`` allocations exported with CUDA IPC handles; the handles travel through the
  ``torch.distributed`` store (the control plane), and each peer maps them
  (``cudaIpcOpenMemHandle``) -- loads/stores to the mapped addresses go over NVLink 5 / NVSwitch.
* **single-process** (in-graph replication, one client driving all GPUs): peer access is enabled
  between the devices and raw pointers are shared directly.

No NCCL call is involved in moving parameters or gradients; NCCL (through ``torch.distributed``)
only bootstraps the store and provides barriers for benchmarking.
"""
from __future__ import annotations

import ctypes
import pickle
import time
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from ..ops import cuda_lib

__all__ = ["Fabric", "FabricBuffer", "view_tensor"]


class _CAI:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def view_tensor(ptr: int, nbytes: int, device: torch.device, dtype: torch.dtype = torch.uint8,
                shape: Optional[Sequence[int]] = None) -> torch.Tensor:
    """Zero-copy torch view of raw device memory (possibly peer-mapped)."""
    with torch.cuda.device(device):
        t = torch.as_tensor(_CAI(ptr, nbytes), device=device)
    t = t.view(dtype)
    return t if shape is None else t.view(*shape)


class FabricBuffer:
    """A named allocation on one rank plus its address as seen from the local process."""

    def __init__(self, name: str, owner: int, ptr: int, nbytes: int, device: torch.device, local: bool):
        self.name, self.owner, self.ptr, self.nbytes, self.device, self.local = name, owner, ptr, nbytes, device, local

    def tensor(self, dtype: torch.dtype = torch.uint8, offset: int = 0, numel: Optional[int] = None) -> torch.Tensor:
        esize = torch.empty((), dtype=dtype).element_size()
        n = (self.nbytes - offset) // esize if numel is None else numel
        return view_tensor(self.ptr + offset, n * esize, self.device, dtype)

    def __repr__(self):
        return "FabricBuffer(%r owner=%d ptr=0x%x bytes=%d local=%s)" % (self.name, self.owner, self.ptr, self.nbytes,
                                                                          self.local)


class Fabric:
    """``local_ranks``: ranks hosted by THIS process, each with its CUDA device ordinal."""

    def __init__(self, world_size: int, local_ranks: Dict[int, int], store=None, prefix: str = "dtf_fabric"):
        self.world_size = world_size
        self.local_ranks = dict(local_ranks)
        self.store = store
        self.prefix = prefix
        self._lib = cuda_lib.load()
        self._owned: Dict[Tuple[int, str], FabricBuffer] = {}
        self._mapped: Dict[Tuple[int, int, str], FabricBuffer] = {}       # (viewer_rank, owner_rank, name)
        self._opened: List[Tuple[int, int]] = []
        self.single_process = len(self.local_ranks) == world_size
        if not self.single_process and store is None:
            raise ValueError("multi-process fabric needs a torch.distributed store for handle exchange")
        if self.single_process and world_size > 1:
            devs = sorted(set(self.local_ranks.values()))
            for d in devs:
                with torch.cuda.device(d):
                    for p in devs:
                        if p != d and self._lib.dtf_can_access_peer(d, p):
                            rc = self._lib.dtf_enable_peer(p)
                            if rc:
                                raise RuntimeError("cudaDeviceEnablePeerAccess(%d->%d) failed: %d" % (d, p, rc))

    @staticmethod
    def from_torch_distributed() -> "Fabric":
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        dev = torch.cuda.current_device()
        store = dist.distributed_c10d._get_default_store()
        return Fabric(world, {rank: dev}, store=store)

    # -- allocation ---------------------------------------------------------------------------------
    def alloc(self, rank: int, name: str, nbytes: int) -> FabricBuffer:
        dev = self.local_ranks[rank]
        nbytes = (int(nbytes) + 255) // 256 * 256
        p = ctypes.c_void_p()
        with torch.cuda.device(dev):
            rc = self._lib.dtf_fabric_alloc(ctypes.byref(p), nbytes)
        if rc:
            raise RuntimeError("fabric alloc of %d bytes on rank %d failed: cudaError %d" % (nbytes, rank, rc))
        buf = FabricBuffer(name, rank, p.value, nbytes, torch.device("cuda", dev), True)
        self._owned[(rank, name)] = buf
        return buf

    # -- publication -----------------------------------------------------------------------------------
    def publish(self, rank: int, name: str) -> None:
        """Export a local buffer so other processes can map it (no-op in single-process mode)."""
        if self.single_process:
            return
        buf = self._owned[(rank, name)]
        h = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(buf.device):
            rc = self._lib.dtf_fabric_export(buf.ptr, h)
        if rc:
            raise RuntimeError("cudaIpcGetMemHandle failed for %s: %d" % (name, rc))
        self.store.set("%s/%d/%s" % (self.prefix, rank, name), pickle.dumps((bytes(h), buf.nbytes)))

    def peer(self, viewer_rank: int, owner_rank: int, name: str, timeout: float = 120.0) -> FabricBuffer:
        """Address of ``owner_rank``'s buffer usable by kernels running on ``viewer_rank``'s GPU."""
        key = (viewer_rank, owner_rank, name)
        got = self._mapped.get(key)
        if got is not None:
            return got
        vdev = self.local_ranks[viewer_rank]
        if owner_rank in self.local_ranks:
            own = self._owned[(owner_rank, name)]
            if own.device.index != vdev and not self.single_process:
                with torch.cuda.device(vdev):
                    self._lib.dtf_enable_peer(own.device.index)
            buf = FabricBuffer(name, owner_rank, own.ptr, own.nbytes, torch.device("cuda", vdev),
                               owner_rank == viewer_rank)
        else:
            skey = "%s/%d/%s" % (self.prefix, owner_rank, name)
            self.store.wait([skey])
            handle, nbytes = pickle.loads(self.store.get(skey))
            hbuf = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
            p = ctypes.c_void_p()
            with torch.cuda.device(vdev):
                rc = self._lib.dtf_fabric_import(hbuf, ctypes.byref(p))
            if rc:
                raise RuntimeError("cudaIpcOpenMemHandle(%s from rank %d) failed: cudaError %d" % (name, owner_rank, rc))
            self._opened.append((vdev, p.value))
            buf = FabricBuffer(name, owner_rank, p.value, nbytes, torch.device("cuda", vdev), False)
        self._mapped[key] = buf
        return buf

    def local(self, rank: int, name: str) -> FabricBuffer:
        return self._owned[(rank, name)]

    def barrier(self) -> None:
        if self.single_process:
            for d in set(self.local_ranks.values()):
                torch.cuda.synchronize(d)
            return
        import torch.distributed as dist
        for d in set(self.local_ranks.values()):
            torch.cuda.synchronize(d)
        dist.barrier()

    def close(self) -> None:
        for dev, p in self._opened:
            with torch.cuda.device(dev):
                self._lib.dtf_fabric_close(p)
        self._opened.clear()
        for buf in self._owned.values():
            with torch.cuda.device(buf.device):
                self._lib.dtf_fabric_free(buf.ptr)
        self._owned.clear()
        self._mapped.clear()
