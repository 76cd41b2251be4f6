"""csrc/fabric_vmm.cu -- UNMODIFIED, compiled by g++ with the real CUDA headers -- against a recording fake of the driver's
VMM / multicast entry points (tests/emu/fake_cuda_driver.cpp, reached through a fake cudaGetDriverEntryPoint): allocation
properties (pinned device memory, POSIX-fd handles), granularity = max(allocation, multicast), reserve -> map -> set-access and
its unwinding when a step fails, fd export / import, multicast create / add-device / bind / unbind argument mapping, and the
support level reported for boxes without VMM or without NVLS."""
import ctypes
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def vmm(tmp_path_factory):
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda.h")):
        pytest.skip("needs g++ and the CUDA toolkit headers")
    so = str(tmp_path_factory.mktemp("emu") / "libvmm_host.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-I" + CUDA_INC, "-x", "c++", "-shared", "-fPIC",
                    "-Wl,-Bsymbolic",        # bind fabric_vmm.cu's cudaGetDriverEntryPoint / cudaFree calls to THIS library's fakes even
                    "-o", so,                # when a real libcudart is already loaded globally in the test process (torch)
                    os.path.join(ROOT, "distributed_tensorflow_b200", "csrc", "fabric_vmm.cu"),
                    os.path.join(ROOT, "tests", "emu", "fake_cuda_driver.cpp")], check=True)
    lib = ctypes.CDLL(so)
    ull, ll, i, vp = ctypes.c_ulonglong, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p
    lib.dtf_vmm_granularity.argtypes = [i, i, ctypes.POINTER(ll)]
    lib.dtf_vmm_create.argtypes = [i, ll, ctypes.POINTER(ull)]
    lib.dtf_vmm_map.argtypes = [ull, ll, i, ctypes.POINTER(vp)]
    lib.dtf_vmm_unmap.argtypes = [vp, ll]
    lib.dtf_vmm_release.argtypes = [ull]
    lib.dtf_vmm_export_fd.argtypes = [ull, ctypes.POINTER(i)]
    lib.dtf_vmm_import_fd.argtypes = [i, ctypes.POINTER(ull)]
    lib.dtf_mc_create.argtypes = [i, ll, ctypes.POINTER(ull)]
    lib.dtf_mc_add_device.argtypes = [ull, i]
    lib.dtf_mc_bind.argtypes = [ull, ull, ll]
    lib.dtf_mc_unbind.argtypes = [ull, i, ll]
    lib.fake_driver_trace.restype = ctypes.c_char_p
    lib.fake_driver_reset.argtypes = [i, i, i, i]
    return lib


def _trace(lib):
    return lib.fake_driver_trace().decode().splitlines()


def test_support_levels_and_granularity(vmm):
    for (v, f, m), want in {(1, 1, 1): 2, (1, 1, 0): 1, (1, 0, 1): 0, (0, 1, 1): 0}.items():
        vmm.fake_driver_reset(0, v, f, m)
        assert vmm.dtf_vmm_support(3) == want
    vmm.fake_driver_reset(0, 1, 1, 1)
    g = ctypes.c_longlong(0)
    assert vmm.dtf_vmm_granularity(2, 8, ctypes.byref(g)) == 0 and g.value == 4 << 20        # max(2 MB allocation, 4 MB multicast)
    assert _trace(vmm) == ["Granularity loc=2 opt=1", "McGranularity ndev=8 opt=1"]
    vmm.fake_driver_reset(0, 1, 1, 1)
    assert vmm.dtf_vmm_granularity(2, 0, ctypes.byref(g)) == 0 and g.value == 2 << 20        # no multicast requested


def test_create_map_export_and_error_unwinding(vmm):
    vmm.fake_driver_reset(0, 1, 1, 1)
    h, p, fd = ctypes.c_ulonglong(0), ctypes.c_void_p(), ctypes.c_int(-1)
    size = 4 << 20
    assert vmm.dtf_vmm_create(5, size, ctypes.byref(h)) == 0 and h.value == 0x1000 + size
    assert vmm.dtf_vmm_map(h.value, size, 5, ctypes.byref(p)) == 0 and p.value == 0x7f0000000000
    assert vmm.dtf_vmm_export_fd(h.value, ctypes.byref(fd)) == 0 and fd.value == (h.value & 0xffff)
    h2 = ctypes.c_ulonglong(0)
    assert vmm.dtf_vmm_import_fd(42, ctypes.byref(h2)) == 0 and h2.value == 0x9000 + 42
    assert vmm.dtf_vmm_unmap(p, size) == 0 and vmm.dtf_vmm_release(h.value) == 0
    assert _trace(vmm) == [
        "MemCreate size=%d type=1 loc=1/5 handle_types=1 flags=0" % size,          # pinned, device 5, POSIX fd exportable
        "AddressReserve size=%d align=0" % size, "Map va=7f0000000000 size=%d off=0 h=%d" % (size, h.value),
        "SetAccess va=7f0000000000 size=%d loc=1/5 flags=3 count=1" % size,        # read-write for device 5
        "Export h=%d type=1" % h.value, "Import fd=42 type=1",
        "Unmap va=7f0000000000 size=%d" % size, "AddressFree va=7f0000000000 size=%d" % size, "MemRelease h=%d" % h.value]
    # the mapping fails at SetAccess (3rd driver call): the reservation and the mapping are undone, the error is reported
    vmm.fake_driver_reset(3, 1, 1, 1)
    assert vmm.dtf_vmm_map(h.value, size, 5, ctypes.byref(p)) == 1                # CUDA_ERROR_INVALID_VALUE
    assert [l.split()[0] for l in _trace(vmm)] == ["AddressReserve", "Map", "SetAccess", "Unmap", "AddressFree"]
    vmm.fake_driver_reset(2, 1, 1, 1)                                              # ... at Map: only the reservation exists
    assert vmm.dtf_vmm_map(h.value, size, 5, ctypes.byref(p)) == 1
    assert [l.split()[0] for l in _trace(vmm)] == ["AddressReserve", "Map", "AddressFree"]


def test_multicast_object_calls(vmm):
    vmm.fake_driver_reset(0, 1, 1, 1)
    mc = ctypes.c_ulonglong(0)
    size = 8 << 20
    assert vmm.dtf_mc_create(8, size, ctypes.byref(mc)) == 0 and mc.value == 0xabc0
    assert vmm.dtf_mc_add_device(mc.value, 6) == 0 and vmm.dtf_mc_bind(mc.value, 0x1234, size) == 0
    assert vmm.dtf_mc_unbind(mc.value, 6, size) == 0
    assert _trace(vmm) == ["McCreate ndev=8 size=%d handle_types=1" % size, "McAddDevice mc=%d dev=6" % mc.value,
                           "McBindMem mc=%d mcoff=0 mem=%d memoff=0 size=%d" % (mc.value, 0x1234, size),
                           "McUnbind mc=%d dev=6 off=0 size=%d" % (mc.value, size)]
