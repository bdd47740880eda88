// Gather-to-root of the final mel shards over NVLink peer memory without an SM-occupying collective kernel
// ("a single gather of the final mel batch").  The reference has no multi-GPU path at all; the
// baseline for this step is ncclAllGather / ncclGather, whose kernel occupies SMs behind the Postnet on the compute stream
// and, as an all-gather, hands every rank (N-1)/N of data it never asked for.  Here:
//
//   * the root rank owns one cudaMalloc'ed receive buffer [flags | shard 0 | shard 1 | ...] and exports it with CUDA IPC;
//   * every other rank maps it (cudaIpcOpenMemHandle) and, once its Postnet has written the step's mel shard, enqueues ONE
//     copy-engine transfer (cudaMemcpyAsync device-to-peer, no SM involved) of its 16 MB shard straight into its slot,
//     followed by a one-thread release-store of the step number into its flag word -- all on a side stream, so the
//     transfer overlaps the next step's kernels;
//   * the root enqueues a one-thread acquire-spin on the flag words where it consumes the batch.
//
// Plain C ABI like the rest of the library; the Python side (fastspeech2_b200/sharded.py::PeerGather) exchanges the
// 64-byte IPC handle through torch.distributed and wraps the raw pointers as tensors.
#include "common.cuh"

namespace fs2 {
namespace {

__global__ void flag_signal_kernel(volatile long long* flag, long long value) {
  __threadfence_system();                 // everything this stream wrote before (incl. the peer copy) is visible system-wide
  *flag = value;
  __threadfence_system();
}

__global__ void flag_wait_kernel(const volatile long long* flags, int n, int skip, long long value) {
  for (int r = (int)threadIdx.x; r < n; r += (int)blockDim.x) {
    if (r == skip) continue;
    unsigned long long spins = 0;
    while (flags[r] < value) {
      __nanosleep(200);
      if (++spins > (1ull << 26)) {       // ~10 s: a peer died or never signalled -- trap instead of hanging the GPU
        printf("fs2 peer gather: rank %d never signalled step %lld (flag = %lld)\n", r, value, (long long)flags[r]);
        __trap();
      }
    }
  }
  __threadfence_system();
}

}  // namespace
}  // namespace fs2

using namespace fs2;

extern "C" {

int fs2_peer_alloc(size_t bytes, void** ptr, void* handle64) {
  FS2_REQUIRE(ptr && handle64 && bytes > 0, "fs2_peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  FS2_CUDA_CHECK(cudaMalloc(&p, bytes));
  FS2_CUDA_CHECK(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); return FS2_ERR_CUDA; }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return FS2_OK;
}

int fs2_peer_free(void* ptr) {
  if (ptr) FS2_CUDA_CHECK(cudaFree(ptr));
  return FS2_OK;
}

int fs2_peer_open(const void* handle64, void** ptr) {
  FS2_REQUIRE(handle64 && ptr, "fs2_peer_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  FS2_CUDA_CHECK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return FS2_OK;
}

int fs2_peer_close(void* ptr) {
  if (ptr) FS2_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
  return FS2_OK;
}

/* dst / src: any two device pointers this process can address (local or peer-mapped); copy engine, asynchronous */
int fs2_peer_copy(void* dst, const void* src, size_t bytes, void* stream) {
  FS2_REQUIRE(dst && src, "fs2_peer_copy: null argument");
  if (bytes == 0) return FS2_OK;
  FS2_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return FS2_OK;
}

int fs2_flag_signal(int64_t* flag, int64_t value, void* stream) {
  FS2_REQUIRE(flag, "fs2_flag_signal: null flag");
  flag_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<volatile long long*>(flag), (long long)value);
  FS2_LAUNCH_CHECK();
  return FS2_OK;
}

int fs2_flag_wait(const int64_t* flags, int n, int skip, int64_t value, void* stream) {
  FS2_REQUIRE(flags && n > 0, "fs2_flag_wait: bad argument");
  flag_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<const volatile long long*>(flags), n, skip, (long long)value);
  FS2_LAUNCH_CHECK();
  return FS2_OK;
}

}  // extern "C"
