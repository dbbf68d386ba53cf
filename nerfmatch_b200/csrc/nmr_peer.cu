// Multi-GPU exchange step of the render path: the rendered feature map and 3D points go to the matcher's
// rank.  Instead of a collective AFTER the render, the compositor's epilogue stores its rows straight into the matcher
// rank's memory: the output pointers of nmr_render_args may be PEER pointers (another GPU's allocation, mapped over
// NVLink / NVSwitch with CUDA IPC), so the transfer rides on the kernel's own coalesced stores, chunk by chunk, while
// the next chunk's MLP runs - no NCCL kernel, no staging copy, nothing left to overlap.  This file holds the plumbing:
// allocate / export / open such buffers, and a release / acquire flag pair so the matcher rank's stream knows when
// every producer's rows of a step have landed.
#include <cstdio>

#include "nmr_internal.cuh"

namespace nmr {

__global__ void peer_signal_kernel(uint32_t* flag, uint32_t value) {
  // runs after every earlier kernel of the stream has completed: their peer stores are performed; publish the step
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

// one thread per producer flag; bounded spin (a lost producer must not hang the box): on timeout the slot of `status`
// (mapped host memory) is set to 1 + the index of the late producer
__global__ void peer_wait_kernel(const uint32_t* flags, int count, uint32_t value, int* status) {
  const int i = threadIdx.x;
  if (i >= count) return;
  const long long t0 = clock64();
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
    if ((int32_t)(v - value) >= 0) break;
    if (clock64() - t0 > 4000000000LL) { atomicExch(status, 1 + i); __threadfence_system(); break; }
    __nanosleep(200);
  }
}

void launch_peer_signal(uint32_t* flag, uint32_t value, cudaStream_t s) { peer_signal_kernel<<<1, 1, 0, s>>>(flag, value); }
void launch_peer_wait(const uint32_t* flags, int count, uint32_t value, int* status, cudaStream_t s) {
  peer_wait_kernel<<<1, 64, 0, s>>>(flags, count, value, status);
}

}  // namespace nmr
