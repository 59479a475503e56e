// afb_peer.cu - the path's collectives as ONE small kernel each over NVLink peer memory.
//
// The only exchanges of the sharded path are tiny: the whole-batch pads (4 floats per rank), d(out)/d(pad) (1 float) and the
// view-parameter gradients ([V, 85] fp32 ~ 2 KB).  Through NCCL each costs a launch plus ~15-35 us of protocol latency,
// which is what limits strong scaling once a rank holds 8 volumes (~0.33 ms of kernels per step).  Here every rank owns a
// buffer in symmetric memory (allocated and exchanged by the host with torch.distributed._symmetric_memory: CUDA VMM
// handles mapped into every peer over NVLink / NVSwitch); one CTA per rank
//     1. PUSHES its contribution into its slot of EVERY rank's buffer (slot = channel x epoch parity x source rank; remote
//        stores over NVLink are fire-and-forget and pipeline, remote loads would cost a round trip each),
//     2. publishes "epoch e is there" into every peer's signal words (st.release.sys after a system fence),
//     3. waits until all peers have published epoch e             (ld.acquire.sys on its OWN signal words, bounded spin),
//     4. reduces the world slots of its OWN buffer in rank order (local reads; bitwise the same result on every rank).
// Two slots per channel (epoch parity) make the buffer reuse race-free without a second barrier: a rank can only push
// epoch e+2 after it has seen every peer publish e+1, which each peer does only after it finished reducing epoch e.
// The epoch counter lives in device memory and is advanced by the kernel itself, so the call is CUDA-graph capturable.
// A peer that never arrives cannot hang the GPU: the spin is bounded (~2 s) and raises a device error flag instead.
#include "afb_device.cuh"

namespace afb {

constexpr int PEER_SIGNAL_WORDS = 256;        // first words of every rank's buffer: signals[channel * world + source rank]
constexpr int PEER_MAX_WORLD = 16;
constexpr int PEER_THREADS = 256;

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// op 0: out[i] = sum_r in_r[i]      op 1: out[r * n + i] = in_r[i] (all-gather)
// op 2: n = 2k floats = k (min, multiplicity) pairs -> whole-batch pairs: out[2i] = min_r in_r[2i], out[2i+1] = sum of the
//       multiplicities of the ranks that hold that minimum (utils/nifti_utils.py:200 under sharding, parallel.allreduce_pad)
// pre_sum > 1: the local contribution is first summed over `pre_sum` rows of `in` (in[k * n + i]): folds the
// d_params.sum(0) over the local batch into the collective.
__global__ void __launch_bounds__(PEER_THREADS)
peer_collective_kernel(float* const* __restrict__ bufs, int rank, int world, int op, int channel, int n, int n_max, int pre_sum,
                       const float* __restrict__ in, float* __restrict__ out, unsigned* __restrict__ epoch, int* __restrict__ err) {
    __shared__ unsigned ep_s;
    if (threadIdx.x == 0) ep_s = epoch[channel] + 1u;
    __syncthreads();
    const unsigned ep = ep_s;
    // slot of source rank r inside any rank's buffer
    const size_t slot0 = (size_t)PEER_SIGNAL_WORDS + ((size_t)channel * 2 + (ep & 1u)) * (size_t)world * (size_t)n_max;
    for (int i = threadIdx.x; i < n; i += PEER_THREADS) {
        float v = in[i];
        for (int k = 1; k < pre_sum; ++k) v += in[(size_t)k * n + i];
        for (int r = 0; r < world; ++r) bufs[r][slot0 + (size_t)rank * n_max + i] = v;      // push to every rank (self included)
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        unsigned* peer_sig = reinterpret_cast<unsigned*>(bufs[threadIdx.x]) + channel * world + rank;
        st_release_sys(peer_sig, ep);
        const unsigned* my_sig = reinterpret_cast<const unsigned*>(bufs[rank]) + channel * world + threadIdx.x;
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(my_sig) - ep) < 0) {
            if (clock64() - t0 > 4000000000ll) { atomicExch(err, 1 + (int)threadIdx.x); break; }
            __nanosleep(20);
        }
    }
    __syncthreads();
    const float* __restrict__ mine = bufs[rank] + slot0;                 // [world][n_max]: what every rank pushed to me
    for (int i = threadIdx.x; i < n; i += PEER_THREADS) {
        if (op == 0) {
            float acc = 0.0f;
            for (int r = 0; r < world; ++r) acc += ld_relaxed_sys(mine + (size_t)r * n_max + i);
            out[i] = acc;
        } else if (op == 2) {
            if ((i & 1) == 0) {
                float m = ld_relaxed_sys(mine + i);
                for (int r = 1; r < world; ++r) m = fminf(m, ld_relaxed_sys(mine + (size_t)r * n_max + i));
                float cnt = 0.0f;
                for (int r = 0; r < world; ++r)
                    if (ld_relaxed_sys(mine + (size_t)r * n_max + i) == m) cnt += ld_relaxed_sys(mine + (size_t)r * n_max + i + 1);
                out[i] = m;
                out[i + 1] = cnt;
            }
        } else {
            for (int r = 0; r < world; ++r) out[(size_t)r * n + i] = ld_relaxed_sys(mine + (size_t)r * n_max + i);
        }
    }
    if (threadIdx.x == 0) epoch[channel] = ep;
}

}  // namespace afb

using namespace afb;

extern "C" int64_t afb_peer_buffer_floats(int n_channels, int n_max, int world) {
    return (int64_t)PEER_SIGNAL_WORDS + (int64_t)n_channels * 2 * world * n_max;
}

/* bufs_dev: device array of `world` pointers, bufs_dev[r] = this process' mapping of rank r's symmetric buffer (each of
 * afb_peer_buffer_floats(n_channels, n_max, world) floats, zero-initialised before the first call).  epoch: device uint32[n_channels],
 * zero-initialised, private to this rank.  err: device int, set non-zero when a peer did not arrive within ~2 s. */
extern "C" int afb_peer_collective(void* const* bufs_dev, int rank, int world, int op, int channel, int n_channels, int n, int n_max,
                                   int pre_sum, const float* in, float* out, void* epoch, int* err, void* stream) {
    if (!bufs_dev || !in || !out || !epoch || !err) return AFB_EINVAL;
    if (world <= 0 || world > PEER_MAX_WORLD || rank < 0 || rank >= world || n <= 0 || n > n_max || pre_sum < 1) return AFB_ESHAPE;
    if (channel < 0 || channel >= n_channels || n_channels * world > PEER_SIGNAL_WORDS || op < 0 || op > 2) return AFB_EINVAL;
    if (op == 2 && (n & 1)) return AFB_ESHAPE;
    peer_collective_kernel<<<1, PEER_THREADS, 0, (cudaStream_t)stream>>>((float* const*)bufs_dev, rank, world, op, channel, n, n_max,
                                                                         pre_sum, in, out, (unsigned*)epoch, err);
    return (int)cudaGetLastError();
}
