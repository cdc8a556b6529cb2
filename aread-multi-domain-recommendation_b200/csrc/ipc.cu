// CUDA IPC plumbing for the row-sharded table: export the allocation that holds a shard, and map a
// peer's shard into this process so that kernels of THIS device can read it over NVLink.  The handle is
// opened with the compute device current and cudaIpcMemLazyEnablePeerAccess, which is what makes the
// mapping visible to this device's kernels (a mapping opened under the owner's device is not).
#include <cuda.h>

#include <cstring>

#include "common.cuh"

namespace aread {
namespace {

using GetRangeFn = CUresult (*)(CUdeviceptr*, size_t*, CUdeviceptr);

GetRangeFn get_range_fn() {
  static GetRangeFn fn = [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &sym, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      sym = nullptr;
    return reinterpret_cast<GetRangeFn>(sym);
  }();
  return fn;
}

}  // namespace
}  // namespace aread

extern "C" {

int aread_ipc_export(const void* ptr, unsigned char* handle_out, int64_t* offset_out) {
  using namespace aread;
  AREAD_REQUIRE(ptr && handle_out && offset_out, "ipc_export: null pointer");
  GetRangeFn fn = get_range_fn();
  if (fn == nullptr) return fail(AREAD_ERR_CUDA, "ipc_export: cuMemGetAddressRange is not available");
  CUdeviceptr base = 0;
  size_t size = 0;
  const CUresult r = fn(&base, &size, reinterpret_cast<CUdeviceptr>(ptr));
  if (r != CUDA_SUCCESS) return fail(AREAD_ERR_CUDA, "ipc_export: cuMemGetAddressRange failed with CUresult %d", (int)r);
  cudaIpcMemHandle_t h;
  AREAD_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
  static_assert(sizeof(h) == AREAD_IPC_HANDLE_BYTES, "unexpected cudaIpcMemHandle_t size");
  std::memcpy(handle_out, &h, sizeof(h));
  *offset_out = static_cast<int64_t>(reinterpret_cast<CUdeviceptr>(ptr) - base);
  return AREAD_OK;
}

int aread_ipc_open(const unsigned char* handle, int64_t offset, int32_t device, void** ptr_out) {
  using namespace aread;
  AREAD_REQUIRE(handle && ptr_out && offset >= 0, "ipc_open: bad arguments");
  AREAD_CUDA(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  void* base = nullptr;
  AREAD_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr_out = static_cast<char*>(base) + offset;
  return AREAD_OK;
}

int aread_ipc_close(void* ptr, int64_t offset) {
  using namespace aread;
  if (ptr == nullptr) return AREAD_OK;
  AREAD_CUDA(cudaIpcCloseMemHandle(static_cast<char*>(ptr) - offset));
  return AREAD_OK;
}

}  // extern "C"
