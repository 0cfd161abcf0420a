// All-reduce (sum) of a small buffer across the GPUs of one NVSwitch domain, done by ONE kernel per rank that
// reads its peers' memory directly (no NCCL call): used by the large-batch data-parallel driver
// (eav_b200/data_parallel.py, BASELINE configs[4]) for the six BatchNorm statistic buffers, the loss and the flat
// 300 KB gradient arena, all of which are latency-bound.
//
// Every rank owns a "symmetric" exchange buffer mapped into all peers (torch symmetric memory; the caller passes the
// table of peer base pointers):   [slot 0 | slot 1 | flags[PR_MAX_CTAS][PR_MAX_WORLD]].
//   1. CTA c copies its chunk of the local source into slot (call & 1) of the OWN exchange buffer,
//   2. publishes flags[c][rank] = call in every peer's buffer (system-scope release) and waits until all peers'
//      CTA c have published the same call number (spin bounded by wall time, then __trap),
//   3. sums chunk c of every rank's slot in RANK ORDER (every rank computes bit-identical results) into dst.
// Two slots make a trailing barrier unnecessary: a rank can be at most one call ahead of a peer that still reads.
#include "eav_common.cuh"
#include "../../include/eav_b200.h"

#ifndef EAV_PEER_TIMEOUT_NS
#define EAV_PEER_TIMEOUT_NS 120000000000ull
#endif

namespace eav {
namespace {

constexpr int PR_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(PR_THREADS)
peer_allreduce_kernel(const T *__restrict__ src, T *__restrict__ dst, int64_t n, const uint64_t *__restrict__ peers,
                      int world, int rank, uint64_t slot_bytes, uint32_t call, const long long *__restrict__ call_base,
                      uint32_t calls_per_step) {
    __shared__ uint64_t base[EAV_PEER_MAX_WORLD];
    if (call_base) call += (uint32_t)(*call_base) * calls_per_step;    // device-resident step counter: graph replays advance
    const int tid = threadIdx.x, cta = blockIdx.x;
    if (tid < world) base[tid] = peers[tid];
    __syncthreads();
    const uint64_t slot_off = (uint64_t)(call & 1u) * slot_bytes;
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)cta * per, hi = lo + per < n ? lo + per : n;
    T *mine = reinterpret_cast<T *>(base[rank] + slot_off);
    for (int64_t i = lo + tid; i < hi; i += PR_THREADS) mine[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (tid < world) {
        const uint64_t flags_off = 2 * slot_bytes + ((uint64_t)cta * EAV_PEER_MAX_WORLD) * sizeof(uint32_t);
        volatile uint32_t *theirs = reinterpret_cast<volatile uint32_t *>(base[tid] + flags_off) + rank;
        *theirs = call;                                           // tell rank `tid` that my chunk `cta` is in place
        volatile uint32_t *ours = reinterpret_cast<volatile uint32_t *>(base[rank] + flags_off) + tid;
        // A late peer (first-step module load, a host stall between eager steps, checkpoint I/O) is waited for, as
        // NCCL would: the bound is wall time (EAV_PEER_TIMEOUT_NS, default 120 s), only a rank that never arrives traps.
        uint32_t spins = 0;
        unsigned long long t0 = 0;
        while ((int32_t)(*ours - call) < 0) {                      // wrap-safe "flag >= call"
            if ((++spins & 0xFFFFu) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > EAV_PEER_TIMEOUT_NS) __trap();
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    for (int64_t i = lo + tid; i < hi; i += PR_THREADS) {
        T s = 0;
        for (int r = 0; r < world; ++r) s += reinterpret_cast<const volatile T *>(base[r] + slot_off)[i];
        dst[i] = s;
    }
}

}  // namespace
}  // namespace eav

extern "C" size_t eav_peer_exchange_bytes(size_t slot_bytes) {
    return 2 * slot_bytes + (size_t)EAV_PEER_MAX_CTAS * EAV_PEER_MAX_WORLD * sizeof(uint32_t);
}

extern "C" int eav_peer_allreduce(const void *src, void *dst, int64_t n, int is_f64, const uint64_t *peer_bases_dev,
                                  int world, int rank, size_t slot_bytes, uint32_t call, const int64_t *call_base_dev,
                                  uint32_t calls_per_step, void *stream) {
    using namespace eav;
    EAV_REQUIRE(src && dst && peer_bases_dev, EAV_ERR_BAD_ARG, "peer_allreduce: null pointer");
    EAV_REQUIRE(world >= 1 && world <= EAV_PEER_MAX_WORLD && rank >= 0 && rank < world, EAV_ERR_BAD_ARG,
                "peer_allreduce: world=%d rank=%d (max world %d)", world, rank, EAV_PEER_MAX_WORLD);
    const size_t esz = is_f64 ? 8 : 4;
    EAV_REQUIRE(n >= 0 && (size_t)n * esz <= slot_bytes && slot_bytes % 16 == 0, EAV_ERR_BAD_ARG,
                "peer_allreduce: %lld elements do not fit the %zu-byte slot", (long long)n, slot_bytes);
    EAV_REQUIRE(call != 0, EAV_ERR_BAD_ARG, "peer_allreduce: call numbers start at 1 (flags are zero-initialised)");
    EAV_REQUIRE(call_base_dev == nullptr || call <= calls_per_step, EAV_ERR_BAD_ARG,
                "peer_allreduce: call %u outside 1..calls_per_step=%u", call, calls_per_step);
    const long long *cb = reinterpret_cast<const long long *>(call_base_dev);
    if (n == 0) return 0;
    int ctas = (int)cdiv64(n, PR_THREADS * 8);
    if (ctas > EAV_PEER_MAX_CTAS) ctas = EAV_PEER_MAX_CTAS;
    cudaStream_t st = (cudaStream_t)stream;
    if (is_f64)
        peer_allreduce_kernel<double><<<ctas, PR_THREADS, 0, st>>>((const double *)src, (double *)dst, n, peer_bases_dev,
                                                                   world, rank, slot_bytes, call, cb, calls_per_step);
    else
        peer_allreduce_kernel<float><<<ctas, PR_THREADS, 0, st>>>((const float *)src, (float *)dst, n, peer_bases_dev, world,
                                                                  rank, slot_bytes, call, cb, calls_per_step);
    EAV_CUDA_LAUNCH_CHECK("peer_allreduce");
    return 0;
}
