// A recording fake of the CUDA driver VMM / multicast entry points, handed to csrc/fabric_vmm.cu through a fake
// cudaGetDriverEntryPoint: the UNMODIFIED fabric_vmm.cu is compiled by g++ with the real CUDA headers and linked against
// this file, so its handle / mapping / error-unwinding logic runs on machines without a GPU (tests/test_fabric_vmm_host.py).
#include <cstdio>
#include <cstring>
#include <string>

#include <cuda.h>
#include <cuda_runtime.h>

namespace {
std::string trace;
int fail_call = 0;          // 1-based index of the call (in trace order) that fails with CUDA_ERROR_INVALID_VALUE
int ncalls = 0;
int attr_vmm = 1, attr_fd = 1, attr_mc = 1;
size_t gran_alloc = 2u << 20, gran_mc = 4u << 20;

template <class... A>
CUresult rec(const char* fmt, A... a) {
  char buf[512];
  std::snprintf(buf, sizeof(buf), fmt, a...);
  trace += buf;
  trace += "\n";
  return (++ncalls == fail_call) ? CUDA_ERROR_INVALID_VALUE : CUDA_SUCCESS;
}

CUresult fMemCreate(CUmemGenericAllocationHandle* h, size_t size, const CUmemAllocationProp* p, unsigned long long flags) {
  *h = 0x1000 + size;
  return rec("MemCreate size=%zu type=%d loc=%d/%d handle_types=%d flags=%llu", size, (int)p->type, (int)p->location.type, p->location.id,
             (int)p->requestedHandleTypes, flags);
}
CUresult fMemRelease(CUmemGenericAllocationHandle h) { return rec("MemRelease h=%llu", (unsigned long long)h); }
CUresult fMemAddressReserve(CUdeviceptr* p, size_t size, size_t align, CUdeviceptr addr, unsigned long long flags) {
  *p = 0x7f0000000000ull;
  return rec("AddressReserve size=%zu align=%zu", size, align);
}
CUresult fMemAddressFree(CUdeviceptr p, size_t size) { return rec("AddressFree va=%llx size=%zu", (unsigned long long)p, size); }
CUresult fMemMap(CUdeviceptr p, size_t size, size_t off, CUmemGenericAllocationHandle h, unsigned long long flags) {
  return rec("Map va=%llx size=%zu off=%zu h=%llu", (unsigned long long)p, size, off, (unsigned long long)h);
}
CUresult fMemUnmap(CUdeviceptr p, size_t size) { return rec("Unmap va=%llx size=%zu", (unsigned long long)p, size); }
CUresult fMemSetAccess(CUdeviceptr p, size_t size, const CUmemAccessDesc* d, size_t count) {
  return rec("SetAccess va=%llx size=%zu loc=%d/%d flags=%d count=%zu", (unsigned long long)p, size, (int)d->location.type, d->location.id,
             (int)d->flags, count);
}
CUresult fMemExport(void* out, CUmemGenericAllocationHandle h, CUmemAllocationHandleType t, unsigned long long flags) {
  *reinterpret_cast<int*>(out) = (int)(h & 0xffff);
  return rec("Export h=%llu type=%d", (unsigned long long)h, (int)t);
}
CUresult fMemImport(CUmemGenericAllocationHandle* h, void* os, CUmemAllocationHandleType t) {
  *h = 0x9000 + (unsigned long long)(uintptr_t)os;
  return rec("Import fd=%llu type=%d", (unsigned long long)(uintptr_t)os, (int)t);
}
CUresult fMemGranularity(size_t* g, const CUmemAllocationProp* p, CUmemAllocationGranularity_flags o) {
  *g = gran_alloc;
  return rec("Granularity loc=%d opt=%d", p->location.id, (int)o);
}
CUresult fMcCreate(CUmemGenericAllocationHandle* h, const CUmulticastObjectProp* p) {
  *h = 0xabc0;
  return rec("McCreate ndev=%u size=%zu handle_types=%llu", p->numDevices, p->size, (unsigned long long)p->handleTypes);
}
CUresult fMcAddDevice(CUmemGenericAllocationHandle h, CUdevice d) { return rec("McAddDevice mc=%llu dev=%d", (unsigned long long)h, (int)d); }
CUresult fMcBindMem(CUmemGenericAllocationHandle mc, size_t mcoff, CUmemGenericAllocationHandle mem, size_t memoff, size_t size,
                    unsigned long long flags) {
  return rec("McBindMem mc=%llu mcoff=%zu mem=%llu memoff=%zu size=%zu", (unsigned long long)mc, mcoff, (unsigned long long)mem, memoff, size);
}
CUresult fMcUnbind(CUmemGenericAllocationHandle mc, CUdevice d, size_t off, size_t size) {
  return rec("McUnbind mc=%llu dev=%d off=%zu size=%zu", (unsigned long long)mc, (int)d, off, size);
}
CUresult fMcGranularity(size_t* g, const CUmulticastObjectProp* p, CUmulticastGranularity_flags o) {
  *g = gran_mc;
  return rec("McGranularity ndev=%u opt=%d", p->numDevices, (int)o);
}
CUresult fDeviceGetAttribute(int* v, CUdevice_attribute a, CUdevice d) {
  *v = a == CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED ? attr_vmm
       : a == CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED ? attr_fd
       : a == CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED ? attr_mc : 0;
  return CUDA_SUCCESS;
}
CUresult fDeviceGet(CUdevice* d, int ordinal) {
  *d = ordinal;
  return CUDA_SUCCESS;
}
}  // namespace

extern "C" {
cudaError_t cudaFree(void*) { return cudaSuccess; }
cudaError_t cudaGetDriverEntryPoint(const char* symbol, void** f, unsigned long long, cudaDriverEntryPointQueryResult* q) {
  struct E { const char* n; void* p; };
  static const E table[] = {
      {"cuMemCreate", (void*)fMemCreate}, {"cuMemRelease", (void*)fMemRelease}, {"cuMemAddressReserve", (void*)fMemAddressReserve},
      {"cuMemAddressFree", (void*)fMemAddressFree}, {"cuMemMap", (void*)fMemMap}, {"cuMemUnmap", (void*)fMemUnmap},
      {"cuMemSetAccess", (void*)fMemSetAccess}, {"cuMemExportToShareableHandle", (void*)fMemExport},
      {"cuMemImportFromShareableHandle", (void*)fMemImport}, {"cuMemGetAllocationGranularity", (void*)fMemGranularity},
      {"cuMulticastCreate", (void*)fMcCreate}, {"cuMulticastAddDevice", (void*)fMcAddDevice}, {"cuMulticastBindMem", (void*)fMcBindMem},
      {"cuMulticastUnbind", (void*)fMcUnbind}, {"cuMulticastGetGranularity", (void*)fMcGranularity},
      {"cuDeviceGetAttribute", (void*)fDeviceGetAttribute}, {"cuDeviceGet", (void*)fDeviceGet}};
  *f = nullptr;
  for (const E& e : table)
    if (std::strcmp(e.n, symbol) == 0) *f = e.p;
  if (q) *q = *f ? cudaDriverEntryPointSuccess : cudaDriverEntryPointSymbolNotFound;
  return cudaSuccess;
}
const char* fake_driver_trace() { return trace.c_str(); }
void fake_driver_reset(int fail_at_call, int vmm, int fd, int mc) {
  trace.clear();
  ncalls = 0;
  fail_call = fail_at_call;
  attr_vmm = vmm;
  attr_fd = fd;
  attr_mc = mc;
}
}
