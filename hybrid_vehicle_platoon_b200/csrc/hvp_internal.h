// hvp_internal.h -- declarations shared by the translation units of libhvp.so (not public).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "miqp_core.cuh"
#include "pm_types.h"

// one device + one stream (include/hvp.h: hvp_ctx)
struct hvp_ctx {
    int device;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    bool timed;
    int64_t launches;
    // grow-only device staging for the *_host entry points
    char* dbuf;
    size_t dcap;
    // pinned host mirror of dbuf for small calls: one H2D + one D2H instead of one copy per array
    char* hbuf;
    size_t hcap;
    // Per-STREAM launch state of the persistent local-MIQP kernel: its work-distribution counter and the scratch rows of
    // the sub-tree adoption (one row per lane of the persistent grid).  Launches on one stream are serialised, so one
    // slot per stream can never be shared by two launches in flight; a stream that finds neither a free nor a reusable
    // slot runs without adoption and with a counter of its own from a ring that is only reused after HVP_COUNTER_RING
    // further launches.
    unsigned long long* counters;                 // [HVP_STREAM_SLOTS + HVP_COUNTER_RING]
    int counter_next;
    cudaStream_t slot_stream[64];
    double* slot_scratch[64];
    int n_slots;
    // When the table is full a NEW stream takes over the least recently used slot whose last launch has FINISHED
    // (slot_done, recorded after every launch outside a capture) -- a caller that creates fresh streams for every call
    // keeps its adoption scratch that way.  A slot that was used while its stream was being captured is pinned: the
    // graph replays its launches later, on streams this table never sees.
    cudaEvent_t slot_done[64];
    unsigned long long slot_use[64];
    bool slot_pinned[64];
    unsigned long long use_clock;
    // side streams of the chunked *_host path (copies of one chunk overlap the kernel of another)
    cudaStream_t side[3];
    cudaEvent_t side_ev;
    bool side_ok;
};

// Timing events of the *_dev entry points.  While the caller's stream is being CAPTURED into a CUDA graph (sweep.py
// replays whole timesteps / ADMM rounds as graphs) a recorded event would belong to the capture and could not be read
// afterwards, so nothing is recorded and the context reports "not timed".
inline cudaError_t hvp_mark(hvp_ctx* c, cudaEvent_t ev, cudaStream_t st, bool last) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) {
        c->timed = false;
        return cudaSuccess;
    }
    const cudaError_t e = cudaEventRecord(ev, st);
    if (last && e == cudaSuccess) c->timed = true;
    return e;
}

// Destroying a handle frees device memory, and a handle is destroyed whenever its owner's garbage collector gets to it --
// possibly while ANOTHER object of the same thread is capturing a CUDA graph (sweep._StepGraph; torch captures in the
// global mode, where a cudaFree anywhere invalidates the capture: seen as a 1-in-3 failure of the GPU suite).  The
// destroy paths therefore run in the relaxed capture mode, as torch's own allocator does for its cudaMalloc calls.
struct HvpRelaxedCapture {
    cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
    HvpRelaxedCapture() { cudaThreadExchangeStreamCaptureMode(&mode); }
    ~HvpRelaxedCapture() { cudaThreadExchangeStreamCaptureMode(&mode); }
};

constexpr int HVP_COUNTER_RING = 256;
constexpr int HVP_STREAM_SLOTS = 64;
constexpr int HVP_MAX_DEVICES = 64;   // per-device caches of function attributes / occupancy
constexpr size_t HVP_STEAL_SLOT_DOUBLES = (size_t)3 << 17;   // 3 MiB of doubles: >= grid x 32 lanes x N for every N (<= 2.6 MB)
int hvp_fail(int code, const char* fmt, ...);          // records the thread's error text, returns code
int hvp_ensure_dbuf(hvp_ctx* c, size_t bytes);

namespace hvp {

constexpr int LOCAL_BLOCK = 32;     // threads (= MIQPs) per CTA of the local-MIQP kernel (one warp)
constexpr int COOP_SPREAD_MAX = 1184;  // batches up to 8 warps per SM: one problem per warp (latency), else packed
constexpr int COOP_BLOCK = 128;     // threads per CTA of the cooperative MIQP kernel (16 groups of 8)
constexpr int FLAT_MIN_BATCH = 8192;  // batches at least this large use the persistent flat kernel (measured crossover with
                                       // the cooperative kernel, N = 6: 5 120 problems 0.86 vs 1.33 ms, 10 240: 2.00 vs 1.68 ms)
constexpr int ROLLOUT_BLOCK = 256;  // threads per CTA of the rollout kernel

// Constants of the rollout kernel (kernel argument; filled by fill_rollout_params).
struct RolloutParams {
    int n, leader_index, flags;
    int scen_per_block;             // whole scenarios handled by one CTA
    double d0, t0, d_safe;
    double tr_t[6][3], tr_v[6][4];  // traction curve (models.py:13-28)
    double inv_rise[6], inv_fall[6]; // RN(1/(v1-v0)), RN(1/(v3-v2)) per gear (exact-division helper)
    double lim[5];                  // gear-switch velocities (models.py:401-403)
    double c_fric, mug;             // friction, mu*g
    double inv_c_fric;              // 1/c_fric (exact when c_fric is a power of two)
    int fric_pow2;                  // c_fric is a power of two (exact-scaling shortcut allowed)
    double default_mass;
};

cudaError_t launch_local_miqp(const LocalParams& P, unsigned long long* counter, double* steal_scratch, int64_t batch, const int32_t* flags, const double* mass,
                              const double* x0, const double* xf, const double* xb, const double* xl,
                              double* u, double* x, int32_t* modes, double* obj, int32_t* status,
                              int32_t* nodes, int32_t* qp_iters, cudaStream_t stream);

cudaError_t launch_fp64_microbench(int iters, double* sink, int* blocks, int* threads, cudaStream_t stream);
cudaError_t launch_smem_microbench(int iters, double* sink, int* blocks, int* threads, cudaStream_t stream);

cudaError_t launch_rollout(const RolloutParams& P, int64_t batch, const double* x, const double* u,
                           const int32_t* gear, const double* mass, const double* leader, double* x_out,
                           double* cost, uint8_t* viol, int32_t* err, cudaStream_t stream);

void pm_layout(PmDev& S);
cudaError_t launch_pm_precompute(const PmDev& S, int64_t batch, const double* x0, const double* params, double* Y,
                                 cudaStream_t stream);
cudaError_t launch_pm_miqp(const PmDev& S, int64_t batch, const double* x0, const double* mass,
                           const double* params, const int32_t* fixed_modes, const double* Y, double* u, double* x,
                           double* extra, int32_t* modes, double* obj, int32_t* status, int32_t* nodes,
                           int32_t* qp_iters, unsigned long long* counter, const PmScratch* scratch, cudaStream_t stream);
cudaError_t launch_pm_shard(const PmDev& S, int64_t batch, const double* x0, const double* mass, const double* params,
                            const double* Y, const double* incumbent, double* u, double* x, double* extra,
                            int32_t* modes, double* obj, int32_t* status, int32_t* nodes, int32_t* qp_iters,
                            unsigned long long* counter, const PmScratch* scratch, cudaStream_t stream);
cudaError_t launch_pm_eval(const PmDev& S, int64_t batch, const double* x0, const double* mass,
                           const double* params, const double* xg, const double* ug, double* cost,
                           cudaStream_t stream);

}  // namespace hvp
