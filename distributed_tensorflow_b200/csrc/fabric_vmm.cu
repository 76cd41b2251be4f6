// Symmetric, multicast-capable device memory for the NVSwitch fabric (CUDA VMM + NVLS multicast objects).
//
// A *symmetric buffer* is one physical allocation per GPU, all the same size, that is
//   * mapped unicast on its owner (ordinary loads/stores),
//   * mapped unicast on every peer (P2P loads/stores over NVLink), and
//   * bound to one multicast object whose mapping accepts `multimem.st` (the switch replicates one store into
//     every GPU's copy -- the ps publishing parameters) and `multimem.ld_reduce` (the switch reads every GPU's
//     copy and returns the sum -- the ps aggregating the workers' gradients).
// Handles travel between processes as POSIX file descriptors (exported here, passed over a unix socket by
// parallel/fdshare.py).  The driver API is resolved at run time (cudaGetDriverEntryPoint), so the library links
// without libcuda and still loads on a machine that has no GPU.
//
// Role in the reference: TF's gRPC RecvTensor transport between /job:ps and /job:worker tasks
// (distributed_mnist.py:74-75,152; SURVEY section 5 "Distributed communication backend").
#include <cstdio>
#include <cstring>
#include <mutex>

#include <cuda.h>
#include <cuda_runtime.h>

namespace {

template <typename F>
F resolve(const char* name) {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  return reinterpret_cast<F>(f);
}

struct Driver {
  decltype(&cuMemCreate) MemCreate;
  decltype(&cuMemRelease) MemRelease;
  decltype(&cuMemAddressReserve) MemAddressReserve;
  decltype(&cuMemAddressFree) MemAddressFree;
  decltype(&cuMemMap) MemMap;
  decltype(&cuMemUnmap) MemUnmap;
  decltype(&cuMemSetAccess) MemSetAccess;
  decltype(&cuMemExportToShareableHandle) MemExport;
  decltype(&cuMemImportFromShareableHandle) MemImport;
  decltype(&cuMemGetAllocationGranularity) MemGranularity;
  decltype(&cuMulticastCreate) McCreate;
  decltype(&cuMulticastAddDevice) McAddDevice;
  decltype(&cuMulticastBindMem) McBindMem;
  decltype(&cuMulticastUnbind) McUnbind;
  decltype(&cuMulticastGetGranularity) McGranularity;
  decltype(&cuDeviceGetAttribute) DeviceGetAttribute;
  decltype(&cuDeviceGet) DeviceGet;
  bool ok = false;
};

Driver& drv() {
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFree(nullptr);      // make sure the runtime (and with it the driver) is initialised
    d.MemCreate = resolve<decltype(&cuMemCreate)>("cuMemCreate");
    d.MemRelease = resolve<decltype(&cuMemRelease)>("cuMemRelease");
    d.MemAddressReserve = resolve<decltype(&cuMemAddressReserve)>("cuMemAddressReserve");
    d.MemAddressFree = resolve<decltype(&cuMemAddressFree)>("cuMemAddressFree");
    d.MemMap = resolve<decltype(&cuMemMap)>("cuMemMap");
    d.MemUnmap = resolve<decltype(&cuMemUnmap)>("cuMemUnmap");
    d.MemSetAccess = resolve<decltype(&cuMemSetAccess)>("cuMemSetAccess");
    d.MemExport = resolve<decltype(&cuMemExportToShareableHandle)>("cuMemExportToShareableHandle");
    d.MemImport = resolve<decltype(&cuMemImportFromShareableHandle)>("cuMemImportFromShareableHandle");
    d.MemGranularity = resolve<decltype(&cuMemGetAllocationGranularity)>("cuMemGetAllocationGranularity");
    d.McCreate = resolve<decltype(&cuMulticastCreate)>("cuMulticastCreate");
    d.McAddDevice = resolve<decltype(&cuMulticastAddDevice)>("cuMulticastAddDevice");
    d.McBindMem = resolve<decltype(&cuMulticastBindMem)>("cuMulticastBindMem");
    d.McUnbind = resolve<decltype(&cuMulticastUnbind)>("cuMulticastUnbind");
    d.McGranularity = resolve<decltype(&cuMulticastGetGranularity)>("cuMulticastGetGranularity");
    d.DeviceGetAttribute = resolve<decltype(&cuDeviceGetAttribute)>("cuDeviceGetAttribute");
    d.DeviceGet = resolve<decltype(&cuDeviceGet)>("cuDeviceGet");
    d.ok = d.MemCreate && d.MemRelease && d.MemAddressReserve && d.MemAddressFree && d.MemMap && d.MemUnmap &&
           d.MemSetAccess && d.MemExport && d.MemImport && d.MemGranularity && d.DeviceGetAttribute && d.DeviceGet;
  });
  return d;
}

CUmemAllocationProp mem_prop(int dev) {
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = dev;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

CUmulticastObjectProp mc_prop(int ndev, size_t size) {
  CUmulticastObjectProp p;
  memset(&p, 0, sizeof(p));
  p.numDevices = (unsigned)ndev;
  p.size = size;
  p.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

}  // namespace

extern "C" {

// 1: VMM allocations with POSIX-fd export work on `dev`; 2: ... and NVLS multicast objects too; 0: neither.
int dtf_vmm_support(int dev) {
  Driver& d = drv();
  if (!d.ok) return 0;
  CUdevice cd;
  if (d.DeviceGet(&cd, dev) != CUDA_SUCCESS) return 0;
  int vmm = 0, fd = 0, mc = 0;
  d.DeviceGetAttribute(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, cd);
  d.DeviceGetAttribute(&fd, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, cd);
  if (!vmm || !fd) return 0;
  if (d.McCreate && d.McAddDevice && d.McBindMem && d.McGranularity)
    d.DeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cd);
  return mc ? 2 : 1;
}

// Size granule a symmetric buffer must be a multiple of (max of the allocation and the multicast granularity).
int dtf_vmm_granularity(int dev, int ndev_multicast, long long* out) {
  Driver& d = drv();
  if (!d.ok) return -1;
  CUmemAllocationProp prop = mem_prop(dev);
  size_t g = 0;
  CUresult r = d.MemGranularity(&g, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
  if (r != CUDA_SUCCESS) return (int)r;
  if (ndev_multicast > 0 && d.McGranularity) {
    CUmulticastObjectProp mp = mc_prop(ndev_multicast, g);
    size_t mg = 0;
    r = d.McGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED);
    if (r == CUDA_SUCCESS && mg > g) g = mg;
  }
  *out = (long long)g;
  return 0;
}

// Physical allocation on `dev` (size must be a multiple of the granularity), exportable as a POSIX fd.
int dtf_vmm_create(int dev, long long size, unsigned long long* handle) {
  Driver& d = drv();
  if (!d.ok) return -1;
  CUmemAllocationProp prop = mem_prop(dev);
  CUmemGenericAllocationHandle h;
  CUresult r = d.MemCreate(&h, (size_t)size, &prop, 0);
  if (r != CUDA_SUCCESS) return (int)r;
  *handle = (unsigned long long)h;
  return 0;
}

// Map `handle` (a memory or a multicast handle) into this process's address space with read/write access for `dev`.
int dtf_vmm_map(unsigned long long handle, long long size, int dev, void** ptr) {
  Driver& d = drv();
  if (!d.ok) return -1;
  CUdeviceptr va = 0;
  CUresult r = d.MemAddressReserve(&va, (size_t)size, 0, 0, 0);
  if (r != CUDA_SUCCESS) return (int)r;
  r = d.MemMap(va, (size_t)size, 0, (CUmemGenericAllocationHandle)handle, 0);
  if (r != CUDA_SUCCESS) {
    d.MemAddressFree(va, (size_t)size);
    return (int)r;
  }
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = dev;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = d.MemSetAccess(va, (size_t)size, &acc, 1);
  if (r != CUDA_SUCCESS) {
    d.MemUnmap(va, (size_t)size);
    d.MemAddressFree(va, (size_t)size);
    return (int)r;
  }
  *ptr = reinterpret_cast<void*>(va);
  return 0;
}

int dtf_vmm_unmap(void* ptr, long long size) {
  Driver& d = drv();
  if (!d.ok) return -1;
  CUresult r = d.MemUnmap(reinterpret_cast<CUdeviceptr>(ptr), (size_t)size);
  CUresult r2 = d.MemAddressFree(reinterpret_cast<CUdeviceptr>(ptr), (size_t)size);
  return (int)(r != CUDA_SUCCESS ? r : r2);
}

int dtf_vmm_release(unsigned long long handle) {
  Driver& d = drv();
  if (!d.ok) return -1;
  return (int)d.MemRelease((CUmemGenericAllocationHandle)handle);
}

int dtf_vmm_export_fd(unsigned long long handle, int* fd) {
  Driver& d = drv();
  if (!d.ok) return -1;
  int out = -1;
  CUresult r = d.MemExport(&out, (CUmemGenericAllocationHandle)handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) return (int)r;
  *fd = out;
  return 0;
}

int dtf_vmm_import_fd(int fd, unsigned long long* handle) {
  Driver& d = drv();
  if (!d.ok) return -1;
  CUmemGenericAllocationHandle h;
  CUresult r = d.MemImport(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
  if (r != CUDA_SUCCESS) return (int)r;
  *handle = (unsigned long long)h;
  return 0;
}

// ---- NVLS multicast objects ---------------------------------------------------------------------------------
int dtf_mc_create(int ndev, long long size, unsigned long long* handle) {
  Driver& d = drv();
  if (!d.ok || !d.McCreate) return -1;
  CUmulticastObjectProp p = mc_prop(ndev, (size_t)size);
  CUmemGenericAllocationHandle h;
  CUresult r = d.McCreate(&h, &p);
  if (r != CUDA_SUCCESS) return (int)r;
  *handle = (unsigned long long)h;
  return 0;
}

// Every participating device must be added before ANY memory is bound.
int dtf_mc_add_device(unsigned long long mc, int dev) {
  Driver& d = drv();
  if (!d.ok || !d.McAddDevice) return -1;
  CUdevice cd;
  CUresult r = d.DeviceGet(&cd, dev);
  if (r != CUDA_SUCCESS) return (int)r;
  return (int)d.McAddDevice((CUmemGenericAllocationHandle)mc, cd);
}

// Bind [0, size) of a device's physical allocation at offset 0 of the multicast object.
int dtf_mc_bind(unsigned long long mc, unsigned long long mem, long long size) {
  Driver& d = drv();
  if (!d.ok || !d.McBindMem) return -1;
  return (int)d.McBindMem((CUmemGenericAllocationHandle)mc, 0, (CUmemGenericAllocationHandle)mem, 0, (size_t)size, 0);
}

int dtf_mc_unbind(unsigned long long mc, int dev, long long size) {
  Driver& d = drv();
  if (!d.ok || !d.McUnbind) return -1;
  CUdevice cd;
  CUresult r = d.DeviceGet(&cd, dev);
  if (r != CUDA_SUCCESS) return (int)r;
  return (int)d.McUnbind((CUmemGenericAllocationHandle)mc, cd, 0, (size_t)size);
}

}  // extern "C"
