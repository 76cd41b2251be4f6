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
import os
import pickle
import time
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from ..ops import cuda_lib

__all__ = ["Fabric", "FabricBuffer", "SymmetricBuffer", "view_tensor"]


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


class SymmetricBuffer:
    """One same-sized VMM allocation per rank, peer-mapped and (when the box supports NVLS) bound to a multicast
    object.  ``local(r)``: rank r's own copy; ``peer(viewer, owner)``: owner's copy as addressable from viewer's
    GPU (NVLink loads/stores); ``mc(viewer)``: the multicast address for viewer's GPU (0 without NVLS) --
    ``multimem.st`` there writes EVERY rank's copy, ``multimem.ld_reduce`` returns the sum over every copy."""

    def __init__(self, fabric: "Fabric", name: str, nbytes: int, size: int):
        self.fabric, self.name, self.nbytes, self.size = fabric, name, nbytes, size
        self.handles: Dict[int, int] = {}            # local rank -> memory handle
        self.local_bufs: Dict[int, FabricBuffer] = {}
        self.peer_bufs: Dict[Tuple[int, int], FabricBuffer] = {}
        self.mc_handle = 0
        self.mc_ptrs: Dict[int, int] = {}            # local rank -> multicast mapping
        self._maps: List[Tuple[int, int]] = []       # (ptr, size) to unmap at close
        self._imported: List[int] = []

    @property
    def multicast(self) -> bool:
        return bool(self.mc_ptrs)

    def local(self, rank: int) -> FabricBuffer:
        return self.local_bufs[rank]

    def mc(self, viewer_rank: int) -> int:
        return self.mc_ptrs.get(viewer_rank, 0)

    def peer(self, viewer_rank: int, owner_rank: int) -> FabricBuffer:
        if viewer_rank == owner_rank:
            return self.local_bufs[owner_rank]
        key = (viewer_rank, owner_rank)
        got = self.peer_bufs.get(key)
        if got is None:
            got = self.fabric._map_symmetric_peer(self, viewer_rank, owner_rank)
            self.peer_bufs[key] = got
        return got


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

    _generation = 0

    @staticmethod
    def from_torch_distributed() -> "Fabric":
        """One rank per process.  Every process must create its fabrics in the same order: the n-th fabric of the
        job gets its own key space in the store, so handles of an earlier (closed) fabric are never picked up."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        dev = torch.cuda.current_device()
        store = dist.distributed_c10d._get_default_store()
        Fabric._generation += 1
        return Fabric(world, {rank: dev}, store=store, prefix="dtf_fabric/g%d" % Fabric._generation)

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

    # -- symmetric (VMM / NVLS multicast) buffers ----------------------------------------------------------
    def _store_min(self, key: str, value: int) -> int:
        """MIN of one integer per process (control plane, through the store)."""
        if self.single_process:
            return value
        me = min(self.local_ranks)
        self.store.set("%s/%s/%d" % (self.prefix, key, me), str(value))
        owners = self._process_leaders()
        keys = ["%s/%s/%d" % (self.prefix, key, o) for o in owners]
        self.store.wait(keys)
        return min(int(self.store.get(k)) for k in keys)

    def _process_leaders(self) -> List[int]:
        """Lowest rank of every process (each process announces its ranks once)."""
        got = getattr(self, "_leaders", None)
        if got is not None:
            return got
        me = min(self.local_ranks)
        for r in self.local_ranks:
            self.store.set("%s/leader_of/%d" % (self.prefix, r), str(me))
        keys = ["%s/leader_of/%d" % (self.prefix, r) for r in range(self.world_size)]
        self.store.wait(keys)
        self._leader_of = {r: int(self.store.get(k)) for r, k in zip(range(self.world_size), keys)}
        self._leaders = sorted(set(self._leader_of.values()))
        return self._leaders

    def _fd_server(self):
        srv = getattr(self, "_fdsrv", None)
        if srv is None:
            from .fdshare import FdServer
            srv = self._fdsrv = FdServer()
            self.store.set("%s/fdsock/%d" % (self.prefix, min(self.local_ranks)), srv.path)
        return srv

    def _fd_path_of(self, owner_rank: int) -> str:
        self._process_leaders()
        key = "%s/fdsock/%d" % (self.prefix, self._leader_of[owner_rank])
        self.store.wait([key])
        return self.store.get(key).decode()

    def nvls_level(self) -> int:
        """0: no VMM export; 1: symmetric peer-mapped buffers only; 2: + NVLS multicast (agreed across processes)."""
        lvl = min(self._lib.dtf_vmm_support(d) for d in self.local_ranks.values())
        return self._store_min("vmm_level", lvl)

    def alloc_symmetric(self, name: str, nbytes: int, multicast: bool = True) -> SymmetricBuffer:
        """COLLECTIVE: every process calls it with the same arguments, in the same order."""
        lib = self._lib
        level = self.nvls_level()
        if level == 0:
            raise RuntimeError("CUDA VMM allocations with POSIX-fd export are not supported on this machine")
        use_mc = bool(multicast) and level == 2 and self.world_size > 1
        dev0 = next(iter(self.local_ranks.values()))
        gran = ctypes.c_longlong(0)
        rc = lib.dtf_vmm_granularity(dev0, self.world_size if use_mc else 0, ctypes.byref(gran))
        if rc:
            raise RuntimeError("cuMemGetAllocationGranularity failed: %d" % rc)
        size = (int(nbytes) + gran.value - 1) // gran.value * gran.value
        sb = SymmetricBuffer(self, name, int(nbytes), size)
        for r, d in self.local_ranks.items():
            h = ctypes.c_ulonglong(0)
            rc = lib.dtf_vmm_create(d, size, ctypes.byref(h))
            if rc:
                raise RuntimeError("cuMemCreate(%d bytes) on rank %d failed: CUresult %d" % (size, r, rc))
            p = ctypes.c_void_p()
            rc = lib.dtf_vmm_map(h.value, size, d, ctypes.byref(p))
            if rc:
                raise RuntimeError("mapping %s on rank %d failed: CUresult %d" % (name, r, rc))
            sb.handles[r] = h.value
            sb._maps.append((p.value, size))
            sb.local_bufs[r] = FabricBuffer(name, r, p.value, size, torch.device("cuda", d), True)
            with torch.cuda.device(d):
                lib.dtf_memset(p.value, 0, size, None)
                torch.cuda.synchronize(d)
            if not self.single_process:
                fd = ctypes.c_int(-1)
                rc = lib.dtf_vmm_export_fd(h.value, ctypes.byref(fd))
                if rc:
                    raise RuntimeError("cuMemExportToShareableHandle failed: %d" % rc)
                self._fd_server().register("%s/mem/%d" % (name, r), fd.value)
        if use_mc:
            mc = ctypes.c_ulonglong(0)
            if self.single_process or 0 in self.local_ranks:
                rc = lib.dtf_mc_create(self.world_size, size, ctypes.byref(mc))
                if rc:
                    raise RuntimeError("cuMulticastCreate failed: CUresult %d" % rc)
                if not self.single_process:
                    fd = ctypes.c_int(-1)
                    rc = lib.dtf_vmm_export_fd(mc.value, ctypes.byref(fd))
                    if rc:
                        raise RuntimeError("export of the multicast handle failed: %d" % rc)
                    self._fd_server().register("%s/mc" % name, fd.value)
            else:
                fd = self._fetch_fd(0, "%s/mc" % name)
                rc = lib.dtf_vmm_import_fd(fd, ctypes.byref(mc))
                os.close(fd)
                if rc:
                    raise RuntimeError("import of the multicast handle failed: CUresult %d" % rc)
            sb.mc_handle = mc.value
            for r, d in self.local_ranks.items():
                rc = lib.dtf_mc_add_device(mc.value, d)
                if rc:
                    raise RuntimeError("cuMulticastAddDevice(dev %d) failed: CUresult %d" % (d, rc))
            self.barrier()                      # every device is in the team before any memory is bound
            for r, d in self.local_ranks.items():
                rc = lib.dtf_mc_bind(mc.value, sb.handles[r], size)
                if rc:
                    raise RuntimeError("cuMulticastBindMem(rank %d) failed: CUresult %d" % (r, rc))
            self.barrier()
            for r, d in self.local_ranks.items():
                p = ctypes.c_void_p()
                rc = lib.dtf_vmm_map(mc.value, size, d, ctypes.byref(p))
                if rc:
                    raise RuntimeError("mapping the multicast object for rank %d failed: CUresult %d" % (r, rc))
                sb.mc_ptrs[r] = p.value
                sb._maps.append((p.value, size))
        self.barrier()
        self._symmetric = getattr(self, "_symmetric", [])
        self._symmetric.append(sb)
        return sb

    def _fetch_fd(self, owner_rank: int, key: str) -> int:
        from .fdshare import fetch_fd
        return fetch_fd(self._fd_path_of(owner_rank), key)

    def _map_symmetric_peer(self, sb: SymmetricBuffer, viewer_rank: int, owner_rank: int) -> FabricBuffer:
        lib = self._lib
        vdev = self.local_ranks[viewer_rank]
        if owner_rank in self.local_ranks:
            handle = sb.handles[owner_rank]
        else:
            fd = self._fetch_fd(owner_rank, "%s/mem/%d" % (sb.name, owner_rank))
            h = ctypes.c_ulonglong(0)
            rc = lib.dtf_vmm_import_fd(fd, ctypes.byref(h))
            os.close(fd)
            if rc:
                raise RuntimeError("import of %s from rank %d failed: CUresult %d" % (sb.name, owner_rank, rc))
            handle = h.value
            sb._imported.append(handle)
        p = ctypes.c_void_p()
        rc = lib.dtf_vmm_map(handle, sb.size, vdev, ctypes.byref(p))
        if rc:
            raise RuntimeError("peer-mapping %s (rank %d -> rank %d) failed: CUresult %d" % (sb.name, owner_rank, viewer_rank, rc))
        sb._maps.append((p.value, sb.size))
        return FabricBuffer(sb.name, owner_rank, p.value, sb.size, torch.device("cuda", vdev), False)

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
        for sb in getattr(self, "_symmetric", []):
            for p, size in sb._maps:
                self._lib.dtf_vmm_unmap(p, size)
            if sb.mc_handle:
                for r, d in self.local_ranks.items():
                    self._lib.dtf_mc_unbind(sb.mc_handle, d, sb.size)
                self._lib.dtf_vmm_release(sb.mc_handle)
            for h in list(sb.handles.values()) + sb._imported:
                self._lib.dtf_vmm_release(h)
            sb._maps, sb.handles, sb._imported, sb.mc_ptrs = [], {}, [], {}
        self._symmetric = []
        srv = getattr(self, "_fdsrv", None)
        if srv is not None:
            srv.close()
            self._fdsrv = None
        for dev, p in self._opened:
            with torch.cuda.device(dev):
                self._lib.dtf_fabric_close(p)
        self._opened.clear()
        for buf in self._owned.values():
            with torch.cuda.device(buf.device):
                self._lib.dtf_fabric_free(buf.ptr)
        self._owned.clear()
        self._mapped.clear()
