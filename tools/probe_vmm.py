"""Probe the driver features the NVLS collectives need (run on the GPU box)."""
from cuda.bindings import driver as cu

def chk(r):
    assert r[0] == cu.CUresult.CUDA_SUCCESS, r[0]
    return r[1] if len(r) == 2 else r[1:]

chk(cu.cuInit(0) + (None,)) if False else cu.cuInit(0)
n = chk(cu.cuDeviceGetCount())
print("devices", n)
A = cu.CUdevice_attribute
for d in range(n):
    dev = chk(cu.cuDeviceGet(d))
    for name in ("CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED", "CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED",
                 "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED", "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED",
                 "CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT", "CU_DEVICE_ATTRIBUTE_COOPERATIVE_LAUNCH"):
        print(d, name, chk(cu.cuDeviceGetAttribute(getattr(A, name), dev)))
