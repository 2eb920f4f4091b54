// All-gather of the packed posterior rows over NVLink peer memory, as ONE kernel per rank.
//
// The MPD regulariser needs every rank's (mu | logvar) rows: 32 KB per rank at the training shapes.  For such
// messages a collective's cost is pure latency (tens of microseconds through a general-purpose library, on the
// critical path of a ~40 us step); here every rank PUSHES its rows straight into each peer's landing buffer with
// 16-byte stores over NVLink, raises a flag in the peer's memory, waits for the peers' flags in its own memory and
// copies the landed slots into a fixed output panel.
//
// Memory (per rank, allocated by mae_comm_alloc and opened by the peers through CUDA IPC):
//   header  : error flag, one launch counter per CTA
//   flags   : [2 parities][world]  uint32, value = epoch in which the slot was written
//   landing : [2 parities][world][slot_floats]
// Epochs live on the device (the kernel is CUDA-graph replayable); consecutive epochs alternate landing parities so a
// fast rank never overwrites data a slow peer is still copying out (a rank can start epoch e+1 only after every peer
// has started epoch e, i.e. finished epoch e-1 entirely).  Spins are bounded: on timeout the kernel sets the error
// flag, NaN-fills its output slot and returns instead of hanging the GPU.
#include "common.cuh"

namespace mae {
namespace {

constexpr int kThreads = 256;
constexpr unsigned long long kTimeoutNs = 2000000000ull;  // 2 s

constexpr int kMaxWorld = 64;
struct CommHeader {        // 512 bytes
    unsigned int error;
    unsigned int pad[63];
    unsigned int epoch[kMaxWorld];   // one counter per CTA (= per peer): no cross-CTA handshake inside the kernel
};

__host__ __device__ inline int64_t flags_off() { return 512; }
__host__ __device__ inline int64_t landing_off(int world) { return 512 + ((2 * world * 4 + 255) / 256) * 256; }

__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// grid = world CTAs; CTA p serves peer p in both directions.
__global__ void __launch_bounds__(kThreads)
p2p_allgather_kernel(const float* __restrict__ local, int64_t n_floats, char* const* __restrict__ peers, int rank, int world,
                     int64_t slot_floats, float* __restrict__ out) {
    __shared__ bool timed_out;
    pdl_launch_dependents();   // the MPD pre-pass behind it may be placed while the flags are awaited
    const int p = blockIdx.x;
    char* own = peers[rank];
    CommHeader* hdr = reinterpret_cast<CommHeader*>(own);
    const unsigned int e = hdr->epoch[p] + 1;   // this CTA's own launch counter (all CTAs of a rank advance in lockstep)
    const int parity = e & 1;
    // ---- push my rows into peer p's landing slot [parity][rank]
    char* dst_base = peers[p];
    float* dst = reinterpret_cast<float*>(dst_base + landing_off(world)) + ((int64_t)parity * world + rank) * slot_floats;
    const int64_t n4 = n_floats / 4;
    const float4* src4 = reinterpret_cast<const float4*>(local);
    float4* dst4 = reinterpret_cast<float4*>(dst);
    for (int64_t i = threadIdx.x; i < n4; i += kThreads) dst4[i] = src4[i];
    for (int64_t i = n4 * 4 + threadIdx.x; i < n_floats; i += kThreads) dst[i] = local[i];
    __syncthreads();                 // the CTA's stores happen-before thread 0's system-scope fence + release
    if (threadIdx.x == 0) {
        __threadfence_system();
        unsigned int* peer_flags = reinterpret_cast<unsigned int*>(dst_base + flags_off());
        st_release_sys(peer_flags + parity * world + rank, e);
        // ---- wait for peer p's rows in my own landing buffer
        const unsigned int* my_flags = reinterpret_cast<const unsigned int*>(own + flags_off());
        const unsigned long long t0 = now_ns();
        bool bad = hdr->error != 0u;   // a previous epoch already timed out: do not spin again
        while (!bad && ld_acquire_sys(my_flags + parity * world + p) != e) {
            if (now_ns() - t0 > kTimeoutNs) {
                bad = true;
                break;
            }
        }
        timed_out = bad;
        if (bad) hdr->error = 1u;
        hdr->epoch[p] = e;
    }
    __syncthreads();
    if (timed_out) {
        // never hand uninitialised rows to the MPD kernels: NaN-fill the slot so the loss of this and every later step is
        // visibly NaN (the error flag is sticky); ShardedHead.check() turns the flag into an exception
        float* o = out + (int64_t)p * n_floats;
        const float nan = __int_as_float(0x7fc00000);
        for (int64_t i = threadIdx.x; i < n_floats; i += kThreads) o[i] = nan;
    } else {
        // ---- copy slot p out of the landing buffer into the fixed output panel (L2 is the coherence point for
        //      the peer's NVLink writes; .cg loads skip this SM's L1)
        const float* land = reinterpret_cast<const float*>(own + landing_off(world)) + ((int64_t)parity * world + p) * slot_floats;
        float* o = out + (int64_t)p * n_floats;
        const float4* l4 = reinterpret_cast<const float4*>(land);
        float4* o4 = reinterpret_cast<float4*>(o);
        for (int64_t i = threadIdx.x; i < n4; i += kThreads) o4[i] = __ldcg(l4 + i);
        for (int64_t i = n4 * 4 + threadIdx.x; i < n_floats; i += kThreads) o[i] = __ldcg(land + i);
    }
}

}  // namespace
}  // namespace mae

extern "C" int64_t mae_comm_buffer_bytes(int64_t world, int64_t slot_floats) {
    if (world < 1 || slot_floats < 1) return 0;
    return mae::landing_off((int)world) + 2 * world * mae::round_up(slot_floats, 4) * 4;
}

extern "C" int mae_comm_alloc(int64_t bytes, void** ptr) {
    MAE_REQUIRE(ptr != nullptr, MAE_ERR_NULL);
    MAE_REQUIRE(bytes > 0, MAE_ERR_SHAPE);
    cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*ptr, 0, (size_t)bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceSynchronize();
    return e == cudaSuccess ? MAE_OK : (int)e;
}

extern "C" int mae_comm_free(void* ptr) {
    cudaError_t e = cudaFree(ptr);
    return e == cudaSuccess ? MAE_OK : (int)e;
}

extern "C" int mae_comm_get_handle(void* ptr, void* handle64) {
    MAE_REQUIRE(ptr && handle64, MAE_ERR_NULL);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaError_t e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), ptr);
    return e == cudaSuccess ? MAE_OK : (int)e;
}

extern "C" int mae_comm_open_handle(const void* handle64, void** ptr) {
    MAE_REQUIRE(ptr && handle64, MAE_ERR_NULL);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    return e == cudaSuccess ? MAE_OK : (int)e;
}

extern "C" int mae_comm_close_handle(void* ptr) {
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    return e == cudaSuccess ? MAE_OK : (int)e;
}

extern "C" int mae_allgather_rows(const float* local, int64_t n_floats, void* const* peer_buffers, int64_t rank, int64_t world,
                                  int64_t slot_floats, float* out, mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(local && peer_buffers && out, MAE_ERR_NULL);
    MAE_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world && n_floats > 0 && n_floats <= slot_floats, MAE_ERR_SHAPE);
    MAE_REQUIRE(slot_floats % 4 == 0, MAE_ERR_ALIGN);
    MAE_REQUIRE((reinterpret_cast<uintptr_t>(local) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                    (n_floats % 4 == 0),
                MAE_ERR_ALIGN);
    p2p_allgather_kernel<<<(unsigned)world, kThreads, 0, as_stream(stream)>>>(
        local, n_floats, reinterpret_cast<char* const*>(peer_buffers), (int)rank, (int)world, slot_floats, out);
    return launch_status();
}

extern "C" int mae_comm_error(const void* own_buffer, int* error_out) {
    MAE_REQUIRE(own_buffer && error_out, MAE_ERR_NULL);
    unsigned int host[1];
    cudaError_t e = cudaMemcpy(host, own_buffer, sizeof(host), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return (int)e;
    *error_out = (int)host[0];
    return MAE_OK;
}
