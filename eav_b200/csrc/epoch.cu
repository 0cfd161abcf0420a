// Device-side control of one training epoch (CNN_torch/EEGNet_tor.py:96-135): the shuffled batch
// schedule the reference gets from DataLoader(shuffle=True) (EEGNet_tor.py:92-93, one fresh permutation
// per epoch), the running per-epoch loss / accuracy sums that train() / validate() print, and the epoch
// counter.  Everything lives on the device so that a WHOLE epoch (9 train steps incl. the ragged last
// batch + the validation pass) is one CUDA graph with no host work between steps.
#include "eegnet_kernels.cuh"

namespace eav {

// One CTA per model.  Thread i draws a 64-bit key from Philox4x32-10 keyed by (seed; subject, epoch, i);
// the permutation is the rank order of the keys (ties broken by index, so it is a bijection).  The stream
// depends on the SUBJECT id, not on the model slot: a subject's batches do not change with the number of
// GPUs the 42 subjects are sharded over.
// sched layout: step s starts at s * M * batch; inside a step [m][b] with B_s = min(batch, n_train - s*batch)
// entries per model, i.e. exactly the x_index vector of eav_eegnet_forward for (M, B_s).
__global__ void __launch_bounds__(256)
epoch_schedule_kernel(int32_t *__restrict__ sched, const int32_t *__restrict__ subject_ids, int n_train, int batch,
                      int M, int64_t rows_per_model, int64_t first_row, uint64_t seed,
                      const long long *__restrict__ epoch_dev) {
    extern __shared__ unsigned long long keys[];
    const int m = blockIdx.x;
    const unsigned long long epoch = epoch_dev ? (unsigned long long)*epoch_dev : 0ull;
    const uint32_t sid = subject_ids ? (uint32_t)subject_ids[m] : (uint32_t)m;
    for (int i = threadIdx.x; i < n_train; i += blockDim.x) {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)i, sid, (uint32_t)epoch, (uint32_t)(epoch >> 32)),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        keys[i] = ((unsigned long long)r.x << 32) | r.y;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_train; i += blockDim.x) {
        const unsigned long long k = keys[i];
        int rank = 0;
        for (int j = 0; j < n_train; ++j) {
            const unsigned long long kj = keys[j];
            rank += (kj < k || (kj == k && j < i)) ? 1 : 0;
        }
        const int s = rank / batch, b = rank - s * batch;
        const int Bs = min(batch, n_train - s * batch);
        sched[(int64_t)s * M * batch + (int64_t)m * Bs + b] = (int32_t)(first_row + (int64_t)m * rows_per_model + i);
    }
}

__global__ void epoch_accumulate_kernel(const float *__restrict__ loss, const int32_t *__restrict__ n_correct, int M,
                                        double *__restrict__ acc) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    acc[2 * m] += (double)loss[m];
    if (n_correct) acc[2 * m + 1] += (double)n_correct[m];
}

__global__ void epoch_commit_kernel(double *__restrict__ train_acc, double *__restrict__ val_acc, int M,
                                    int n_train_steps, int n_val_steps, int n_val, float *__restrict__ history,
                                    int max_epochs, long long *__restrict__ epoch_dev) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const long long e = *epoch_dev;
    if (m < M) {
        float *h = history + ((e % max_epochs) * (long long)M + m) * 3;
        h[0] = n_train_steps > 0 ? (float)(train_acc[2 * m] / n_train_steps) : 0.f;
        train_acc[2 * m] = train_acc[2 * m + 1] = 0.0;
        if (val_acc != nullptr) {
            h[1] = n_val_steps > 0 ? (float)(val_acc[2 * m] / n_val_steps) : 0.f;
            h[2] = n_val > 0 ? (float)(val_acc[2 * m + 1] / n_val) : 0.f;
            val_acc[2 * m] = val_acc[2 * m + 1] = 0.0;
        }
    }
    // every thread has read the counter before any thread of this (single-CTA) launch bumps it
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) *epoch_dev = e + 1;
}

// validation results of epoch (*epoch_dev + epoch_offset), committed one graph later than its training loss when the
// validation pass is pipelined with the next epoch's training (trainer_core.EpochRunner)
__global__ void epoch_commit_val_kernel(double *__restrict__ val_acc, int M, int n_val_steps, int n_val,
                                        float *__restrict__ history, int max_epochs, const long long *__restrict__ epoch_dev,
                                        int epoch_offset) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    long long e = *epoch_dev + epoch_offset;
    if (e < 0) e = 0;
    float *h = history + ((e % max_epochs) * (long long)M + m) * 3;
    h[1] = n_val_steps > 0 ? (float)(val_acc[2 * m] / n_val_steps) : 0.f;
    h[2] = n_val > 0 ? (float)(val_acc[2 * m + 1] / n_val) : 0.f;
    val_acc[2 * m] = val_acc[2 * m + 1] = 0.0;
}

}  // namespace eav

using namespace eav;

extern "C" int eav_epoch_schedule(int32_t *sched, const int32_t *subject_ids, int32_t n_models, int32_t n_train,
                                  int32_t batch, int64_t rows_per_model, int64_t first_row, uint64_t seed,
                                  const int64_t *epoch_dev, void *stream) {
    EAV_REQUIRE(sched != nullptr, EAV_ERR_BAD_ARG, "epoch_schedule: null schedule buffer");
    EAV_REQUIRE(n_models > 0 && n_train > 0 && batch > 0, EAV_ERR_BAD_ARG,
                "epoch_schedule: n_models=%d n_train=%d batch=%d must be positive", n_models, n_train, batch);
    EAV_REQUIRE(n_train <= 4096, EAV_ERR_UNSUPPORTED, "epoch_schedule: n_train=%d > 4096", n_train);
    EAV_REQUIRE(first_row >= 0 && rows_per_model >= 0 &&
                    first_row + (int64_t)(n_models - 1) * rows_per_model + n_train <= INT32_MAX,
                EAV_ERR_BAD_ARG, "epoch_schedule: row numbers do not fit int32");
    epoch_schedule_kernel<<<n_models, 256, (size_t)n_train * sizeof(unsigned long long), (cudaStream_t)stream>>>(
        sched, subject_ids, n_train, batch, n_models, rows_per_model, first_row, seed,
        reinterpret_cast<const long long *>(epoch_dev));
    EAV_CUDA_LAUNCH_CHECK("epoch_schedule");
    return 0;
}

extern "C" int eav_epoch_accumulate(const float *loss, const int32_t *n_correct, int32_t n_models, double *acc,
                                    void *stream) {
    EAV_REQUIRE(loss && acc && n_models > 0, EAV_ERR_BAD_ARG, "epoch_accumulate: bad argument");
    epoch_accumulate_kernel<<<cdiv(n_models, 128), 128, 0, (cudaStream_t)stream>>>(loss, n_correct, n_models, acc);
    EAV_CUDA_LAUNCH_CHECK("epoch_accumulate");
    return 0;
}

extern "C" int eav_epoch_commit(double *train_acc, double *val_acc, int32_t n_models, int32_t n_train_steps,
                                int32_t n_val_steps, int32_t n_val, float *history, int32_t max_epochs,
                                int64_t *epoch_dev, void *stream) {
    EAV_REQUIRE(train_acc && history && epoch_dev, EAV_ERR_BAD_ARG, "epoch_commit: null pointer");
    EAV_REQUIRE(n_models > 0 && n_models <= 1024 && max_epochs > 0, EAV_ERR_BAD_ARG,
                "epoch_commit: n_models=%d (<= 1024) max_epochs=%d", n_models, max_epochs);
    epoch_commit_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(train_acc, val_acc, n_models, n_train_steps, n_val_steps,
                                                             n_val, history, max_epochs,
                                                             reinterpret_cast<long long *>(epoch_dev));
    EAV_CUDA_LAUNCH_CHECK("epoch_commit");
    return 0;
}

extern "C" int eav_epoch_commit_val(double *val_acc, int32_t n_models, int32_t n_val_steps, int32_t n_val, float *history,
                                    int32_t max_epochs, const int64_t *epoch_dev, int32_t epoch_offset, void *stream) {
    EAV_REQUIRE(val_acc && history && epoch_dev, EAV_ERR_BAD_ARG, "epoch_commit_val: null pointer");
    EAV_REQUIRE(n_models > 0 && max_epochs > 0, EAV_ERR_BAD_ARG, "epoch_commit_val: n_models=%d max_epochs=%d", n_models, max_epochs);
    epoch_commit_val_kernel<<<cdiv(n_models, 128), 128, 0, (cudaStream_t)stream>>>(
        val_acc, n_models, n_val_steps, n_val, history, max_epochs, reinterpret_cast<const long long *>(epoch_dev), epoch_offset);
    EAV_CUDA_LAUNCH_CHECK("epoch_commit_val");
    return 0;
}
