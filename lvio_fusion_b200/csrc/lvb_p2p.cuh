// lvb_p2p.cuh -- the in-kernel all-reduce over NVLink peer memory, as device functions: ctx.cu wraps them in the generic
// p2p_allreduce_kernel, ba.cu fuses them with the scalar pack / unpack and the camera damping of an LM pass (one kernel instead of four).
#pragma once
#include "lvb_internal.cuh"

namespace lvb {

// ---- in-kernel all-reduce over peer memory (NVLink / NVSwitch) ---------------------------------------------------------
// The reduced system of a window is ~0.2 MB: an NCCL call costs more in launch + protocol latency than the transfer.  Each
// rank owns an exchange buffer (cudaMalloc + CUDA IPC, mapped by every peer).  One kernel per all-reduce:
//   1. every CTA copies its slice of the local contribution into the local exchange buffer (phase = epoch parity),
//   2. release-stores the epoch into the flag slot (phase, CTA, my rank) of every peer,
//   3. acquire-spins on its own flag slots until all peers have published the same epoch,
//   4. sums the slice over the ranks IN RANK ORDER with loads from the peers' buffers (bitwise identical result on every
//      rank, which the redundant solves rely on) and writes it back in place.
// The phase double-buffers the data: a rank can only reach epoch e+2 after every peer has published e+1, i.e. after they
// finished reading e.  The epoch lives on the device and is advanced by the kernel, so the launch can sit in a CUDA graph.
enum { XB_BLOCKS = 64, XB_DATA = 1 << 20, XB_FLAGS = 2 * XB_BLOCKS * 8 * 4, XB_CTRL = 256, XB_TOTAL = XB_FLAGS + XB_CTRL + 2 * XB_DATA };
struct P2PArgs { unsigned char* peer[8]; int rank, world; unsigned long long timeout_ns; };

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) { unsigned int v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// The exchange proper.  `load(i)` yields this rank's contribution for element i (the fused BA kernel substitutes the packed scalars
// there); returns false when the communicator is dead (sticky error word) -- the caller must then skip its epilogue.  On return
// buf holds the sum over the ranks for this CTA's elements; `last` tells the one CTA that finished last (all CTAs' results are
// then visible to it): it may run a grid-wide epilogue, after which it must call p2p_finish_epoch.
template <class Load>
__device__ __forceinline__ bool p2p_allreduce_body(const P2PArgs& a, double* __restrict__ buf, int count, Load load, bool* last, unsigned int* epoch_out) {
    unsigned char* mine = a.peer[a.rank];
    unsigned int* ctrl = reinterpret_cast<unsigned int*>(mine + XB_FLAGS);          // [0] epoch, [1] finished CTAs, [2] sticky error (1 + silent peer)
    __shared__ int s_dead;
    if (threadIdx.x == 0) s_dead = *reinterpret_cast<volatile unsigned int*>(ctrl + 2) != 0u;
    __syncthreads();
    if (s_dead) return false;                   // an earlier exchange timed out: the communicator is dead, the host reports LVB_ERR_COMM
    const unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(ctrl) + 1u;
    const int phase = (int)(epoch & 1u);
    double* my_data = reinterpret_cast<double*>(mine + XB_FLAGS + XB_CTRL + (size_t)phase * XB_DATA);
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) my_data[i] = load(i);
    __syncthreads();
    const size_t slot = ((size_t)phase * XB_BLOCKS + blockIdx.x) * 8;
    if (threadIdx.x < a.world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned int*>(a.peer[threadIdx.x]) + slot + a.rank, epoch);
        const unsigned int* f = reinterpret_cast<const unsigned int*>(mine) + slot + threadIdx.x;
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (ld_acquire_sys(f) != epoch) {
            // a lost or mismatched peer must neither hang the GPU nor kill the process: after the timeout the kernel records which
            // peer stayed silent and returns; every later exchange returns at once and the host turns the flag into LVB_ERR_COMM
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > a.timeout_ns) { atomicCAS(ctrl + 2, 0u, 1u + (unsigned int)threadIdx.x); s_dead = 1; break; }
        }
    }
    __syncthreads();
    if (s_dead) return false;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        double s = 0.0;
        for (int r = 0; r < a.world; ++r) s += __ldcv(reinterpret_cast<const double*>(a.peer[r] + XB_FLAGS + XB_CTRL + (size_t)phase * XB_DATA) + i);
        buf[i] = s;
    }
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(&ctrl[1], 1u) == gridDim.x - 1) ? 1 : 0;
        if (s_last) __threadfence();
    }
    __syncthreads();
    *last = s_last != 0;
    *epoch_out = epoch;
    return true;
}
// by the last CTA, once its epilogue (if any) is done: the next exchange may start
__device__ __forceinline__ void p2p_finish_epoch(const P2PArgs& a, unsigned int epoch) {
    unsigned int* ctrl = reinterpret_cast<unsigned int*>(a.peer[a.rank] + XB_FLAGS);
    if (threadIdx.x == 0) { ctrl[1] = 0u; __threadfence(); *reinterpret_cast<volatile unsigned int*>(ctrl) = epoch; }
}

bool comm_p2p_args(const lvb_ctx* ctx, size_t count, P2PArgs* a);      // ctx.cu

}  // namespace lvb

