// Symmetric memory for the NVLink / NVSwitch collectives: CUDA virtual-memory-management allocations that are
//   (a) mapped into every peer process (unicast P2P pointers: ld/st over NVLink), and
//   (b) bound to an NVSwitch MULTICAST object, whose mapping gives one address that
//       - `multimem.st`        replicates a store into every rank's copy inside the switch, and
//       - `multimem.ld_reduce` returns the sum over every rank's copy, reduced inside the switch (NVLS).
//
// One "segment" = one physical allocation per rank of the same size, created collectively:
//   every rank : cuMemCreate(POSIX-fd shareable) -> export fd            (fl4h_symm_create)
//   (python)   : fds travel between the rank processes over AF_UNIX / SCM_RIGHTS
//   every rank : import + map each peer's allocation                      (fl4h_symm_map_peer)
//   rank 0     : cuMulticastCreate -> export fd                           (fl4h_symm_mc_create)
//   every rank : import, cuMulticastAddDevice                             (fl4h_symm_mc_join)
//                --- barrier: all devices added ---
//                cuMulticastBindMem(own allocation), map the multicast VA (fl4h_symm_mc_bind)
//
// Driver entry points are resolved at run time (cudaGetDriverEntryPoint): the library still loads on driver-less build
// machines.  Replaces the reference's transport for the same bytes: gRPC messages of np.save blobs
// (fl4health/parameter_exchange/full_exchanger.py:30,45-47).

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <unistd.h>

namespace {

template <typename Fn>
Fn driver_fn(const char* name) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult status;
    if (cudaGetDriverEntryPoint(name, &ptr, cudaEnableDefault, &status) != cudaSuccess ||
        status != cudaDriverEntryPointSuccess) {
        return nullptr;
    }
    return reinterpret_cast<Fn>(ptr);
}

#define DRV(name, type)                                    \
    static type fn_##name = driver_fn<type>(#name);        \
    if (fn_##name == nullptr) return (int)CUDA_ERROR_NOT_SUPPORTED;

using MemCreateFn = CUresult (*)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
using MemReleaseFn = CUresult (*)(CUmemGenericAllocationHandle);
using MemExportFn = CUresult (*)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
using MemImportFn = CUresult (*)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
using MemReserveFn = CUresult (*)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
using MemMapFn = CUresult (*)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
using MemUnmapFn = CUresult (*)(CUdeviceptr, size_t);
using MemFreeVaFn = CUresult (*)(CUdeviceptr, size_t);
using MemSetAccessFn = CUresult (*)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
using MemGranFn = CUresult (*)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
using McCreateFn = CUresult (*)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
using McAddDevFn = CUresult (*)(CUmemGenericAllocationHandle, CUdevice);
using McBindMemFn = CUresult (*)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t,
                                 unsigned long long);
using McGranFn = CUresult (*)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
using DevGetAttrFn = CUresult (*)(int*, CUdevice_attribute, CUdevice);

CUmemAllocationProp device_prop(int dev) {
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = dev;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return prop;
}

CUmulticastObjectProp mc_prop(int world, size_t bytes) {
    CUmulticastObjectProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.numDevices = (unsigned)world;
    prop.size = bytes;
    prop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return prop;
}

int map_rw(CUmemGenericAllocationHandle handle, size_t bytes, size_t align, int dev, void** out) {
    DRV(cuMemAddressReserve, MemReserveFn);
    DRV(cuMemMap, MemMapFn);
    DRV(cuMemSetAccess, MemSetAccessFn);
    DRV(cuMemAddressFree, MemFreeVaFn);
    CUdeviceptr va = 0;
    CUresult err = fn_cuMemAddressReserve(&va, bytes, align, 0, 0);
    if (err != CUDA_SUCCESS) return (int)err;
    err = fn_cuMemMap(va, bytes, 0, handle, 0);
    if (err != CUDA_SUCCESS) {
        fn_cuMemAddressFree(va, bytes);
        return (int)err;
    }
    CUmemAccessDesc access;
    memset(&access, 0, sizeof(access));
    access.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    access.location.id = dev;
    access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    err = fn_cuMemSetAccess(va, bytes, &access, 1);
    if (err != CUDA_SUCCESS) return (int)err;
    *out = reinterpret_cast<void*>(va);
    return 0;
}

}  // namespace

extern "C" {

// bit 0: VMM, bit 1: POSIX-fd handles, bit 2: multicast (NVLS)
int fl4h_symm_features(int dev) {
    static DevGetAttrFn get = driver_fn<DevGetAttrFn>("cuDeviceGetAttribute");
    if (get == nullptr) return 0;
    cudaFree(nullptr);  // make sure the primary context exists
    int vmm = 0, fd = 0, mc = 0;
    get(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev);
    get(&fd, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev);
    get(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
    return (vmm ? 1 : 0) | (fd ? 2 : 0) | (mc ? 4 : 0);
}

// Size every rank must use for a segment of `bytes`: a multiple of the allocation and (if used) multicast granularity.
int fl4h_symm_round_size(int dev, int world, int use_multicast, size_t bytes, size_t* rounded) {
    DRV(cuMemGetAllocationGranularity, MemGranFn);
    CUmemAllocationProp prop = device_prop(dev);
    size_t gran = 0;
    CUresult err = fn_cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    if (err != CUDA_SUCCESS) return (int)err;
    if (use_multicast) {
        DRV(cuMulticastGetGranularity, McGranFn);
        CUmulticastObjectProp mprop = mc_prop(world, bytes);
        size_t mgran = 0;
        err = fn_cuMulticastGetGranularity(&mgran, &mprop, CU_MULTICAST_GRANULARITY_RECOMMENDED);
        if (err != CUDA_SUCCESS) return (int)err;
        if (mgran > gran) gran = mgran;
    }
    *rounded = (bytes + gran - 1) / gran * gran;
    return 0;
}

// Physical allocation on `dev` (zero-filled), mapped read/write for `dev`; `fd_out` is the shareable handle.
int fl4h_symm_create(int dev, size_t bytes, unsigned long long* handle_out, void** ptr_out, int* fd_out) {
    DRV(cuMemCreate, MemCreateFn);
    DRV(cuMemExportToShareableHandle, MemExportFn);
    cudaFree(nullptr);
    CUmemAllocationProp prop = device_prop(dev);
    CUmemGenericAllocationHandle handle = 0;
    CUresult err = fn_cuMemCreate(&handle, bytes, &prop, 0);
    if (err != CUDA_SUCCESS) return (int)err;
    int rc = map_rw(handle, bytes, 0, dev, ptr_out);
    if (rc != 0) return rc;
    cudaError_t rerr = cudaMemset(*ptr_out, 0, bytes);
    if (rerr != cudaSuccess) return (int)rerr;
    rerr = cudaDeviceSynchronize();
    if (rerr != cudaSuccess) return (int)rerr;
    int fd = -1;
    err = fn_cuMemExportToShareableHandle(&fd, handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    if (err != CUDA_SUCCESS) return (int)err;
    *handle_out = (unsigned long long)handle;
    *fd_out = fd;
    return 0;
}

// Import a peer's allocation from its fd and map it read/write for `dev` (P2P over NVLink).  The fd is closed.
int fl4h_symm_map_peer(int dev, int fd, size_t bytes, unsigned long long* handle_out, void** ptr_out) {
    DRV(cuMemImportFromShareableHandle, MemImportFn);
    CUmemGenericAllocationHandle handle = 0;
    CUresult err = fn_cuMemImportFromShareableHandle(&handle, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    close(fd);
    if (err != CUDA_SUCCESS) return (int)err;
    *handle_out = (unsigned long long)handle;
    return map_rw(handle, bytes, 0, dev, ptr_out);
}

// Rank 0: create the multicast object for `world` devices; `fd_out` goes to every other rank.
int fl4h_symm_mc_create(int world, size_t bytes, unsigned long long* mc_out, int* fd_out) {
    DRV(cuMulticastCreate, McCreateFn);
    DRV(cuMemExportToShareableHandle, MemExportFn);
    CUmulticastObjectProp prop = mc_prop(world, bytes);
    CUmemGenericAllocationHandle mc = 0;
    CUresult err = fn_cuMulticastCreate(&mc, &prop);
    if (err != CUDA_SUCCESS) return (int)err;
    int fd = -1;
    err = fn_cuMemExportToShareableHandle(&fd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    if (err != CUDA_SUCCESS) return (int)err;
    *mc_out = (unsigned long long)mc;
    *fd_out = fd;
    return 0;
}

// Other ranks: import the multicast object (fd < 0: `mc_inout` already holds rank 0's own handle); add this device.
int fl4h_symm_mc_join(int dev, int fd, unsigned long long* mc_inout) {
    DRV(cuMulticastAddDevice, McAddDevFn);
    if (fd >= 0) {
        DRV(cuMemImportFromShareableHandle, MemImportFn);
        CUmemGenericAllocationHandle mc = 0;
        CUresult err = fn_cuMemImportFromShareableHandle(&mc, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
        close(fd);
        if (err != CUDA_SUCCESS) return (int)err;
        *mc_inout = (unsigned long long)mc;
    }
    return (int)fn_cuMulticastAddDevice((CUmemGenericAllocationHandle)*mc_inout, (CUdevice)dev);
}

// After EVERY rank joined: bind this rank's allocation at offset 0 and map the multicast address range.
int fl4h_symm_mc_bind(int dev, unsigned long long mc, unsigned long long mem_handle, size_t bytes, void** mc_ptr_out) {
    DRV(cuMulticastBindMem, McBindMemFn);
    CUresult err = fn_cuMulticastBindMem((CUmemGenericAllocationHandle)mc, 0, (CUmemGenericAllocationHandle)mem_handle, 0,
                                         bytes, 0);
    if (err != CUDA_SUCCESS) return (int)err;
    return map_rw((CUmemGenericAllocationHandle)mc, bytes, 0, dev, mc_ptr_out);
}

int fl4h_symm_unmap(void* ptr, size_t bytes, unsigned long long handle) {
    DRV(cuMemUnmap, MemUnmapFn);
    DRV(cuMemAddressFree, MemFreeVaFn);
    DRV(cuMemRelease, MemReleaseFn);
    CUresult err = CUDA_SUCCESS;
    if (ptr != nullptr) {
        err = fn_cuMemUnmap((CUdeviceptr)ptr, bytes);
        if (err == CUDA_SUCCESS) err = fn_cuMemAddressFree((CUdeviceptr)ptr, bytes);
    }
    if (handle != 0) {
        CUresult rel = fn_cuMemRelease((CUmemGenericAllocationHandle)handle);
        if (err == CUDA_SUCCESS) err = rel;
    }
    return (int)err;
}

}  // extern "C"
