// hvp_api.cu -- the C ABI of libhvp.so (include/hvp.h): contexts, error reporting, host-buffer
// wrappers.  No torch types, no CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/hvp.h"
#include "hvp_internal.h"
#include "vehicle_model.h"

using namespace hvp;

static thread_local char g_err[512] = "";

int hvp_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define fail hvp_fail
#define CUDA_TRY(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fail(-100 - (int)e__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                   \
    } while (0)


extern "C" int hvp_version(void) { return HVP_VERSION; }

extern "C" int hvp_last_error(char* buf, int len) {
    if (buf && len > 0) {
        strncpy(buf, g_err, (size_t)len - 1);
        buf[len - 1] = 0;
    }
    return (int)strlen(g_err);
}

extern "C" int hvp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int hvp_ctx_create(int device, hvp_ctx** out) {
    if (!out) return fail(-1, "hvp_ctx_create: out is NULL");
    int n = hvp_device_count();
    if (n <= 0) return fail(-2, "hvp_ctx_create: no CUDA device visible (this library has no CPU path)");
    if (device < 0 || device >= n) return fail(-3, "hvp_ctx_create: device %d out of range [0,%d)", device, n);
    CUDA_TRY(cudaSetDevice(device));
    hvp_ctx* c = new hvp_ctx();
    c->device = device; c->timed = false; c->launches = 0; c->dbuf = nullptr; c->dcap = 0; c->hbuf = nullptr; c->hcap = 0;
    c->counters = nullptr; c->counter_next = 0; c->side_ok = false; c->n_slots = 0; c->use_clock = 0;
    c->stream = nullptr; c->ev0 = nullptr; c->ev1 = nullptr;
    // a failure half-way must not leak what was created before it
    cudaError_t e = cudaMalloc(&c->counters, (HVP_STREAM_SLOTS + HVP_COUNTER_RING) * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e != cudaSuccess) {
        if (c->ev1) cudaEventDestroy(c->ev1);
        if (c->ev0) cudaEventDestroy(c->ev0);
        if (c->stream) cudaStreamDestroy(c->stream);
        if (c->counters) cudaFree(c->counters);
        delete c;
        return fail(-100 - (int)e, "hvp_ctx_create: %s", cudaGetErrorString(e));
    }
    *out = c;
    return 0;
}

extern "C" int hvp_ctx_destroy(hvp_ctx* c) {
    if (!c) return 0;
    HvpRelaxedCapture relaxed__;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->dbuf) cudaFree(c->dbuf);
    if (c->hbuf) cudaFreeHost(c->hbuf);
    if (c->counters) cudaFree(c->counters);
    for (int i = 0; i < c->n_slots; ++i) { if (c->slot_scratch[i]) cudaFree(c->slot_scratch[i]); cudaEventDestroy(c->slot_done[i]); }
    if (c->side_ok) { for (int i = 0; i < 3; ++i) cudaStreamDestroy(c->side[i]); cudaEventDestroy(c->side_ev); }
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

extern "C" void* hvp_ctx_stream(hvp_ctx* c) { return c ? (void*)c->stream : nullptr; }

extern "C" int hvp_ctx_synchronize(hvp_ctx* c) {
    if (!c) return fail(-1, "ctx is NULL");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int64_t hvp_ctx_launch_count(hvp_ctx* c) { return c ? c->launches : 0; }

extern "C" float hvp_ctx_last_kernel_ms(hvp_ctx* c) {
    if (!c || !c->timed) return -1.f;
    if (cudaEventSynchronize(c->ev1) != cudaSuccess) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return -1.f;
    return ms;
}

int hvp_ensure_dbuf(hvp_ctx* c, size_t bytes) {
    if (bytes <= c->dcap) return 0;
    if (c->dbuf) { CUDA_TRY(cudaStreamSynchronize(c->stream)); CUDA_TRY(cudaFree(c->dbuf)); c->dbuf = nullptr; c->dcap = 0; }
    size_t cap = bytes + bytes / 4 + 4096;
    CUDA_TRY(cudaMalloc(&c->dbuf, cap));
    c->dcap = cap;
    return 0;
}

#define ensure_dbuf hvp_ensure_dbuf
static inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

// ------------------------------------------------------------------------------------------
// rollout
// ------------------------------------------------------------------------------------------
static int fill_rollout_params(RolloutParams& P, const hvp_env_desc* d) {
    if (!d) return fail(-1, "rollout: desc is NULL");
    if (d->n < 1 || d->n > ROLLOUT_BLOCK) return fail(-4, "rollout: n=%d out of range [1,%d]", d->n, ROLLOUT_BLOCK);
    if (d->leader_index < 0 || d->leader_index >= d->n)
        return fail(-4, "rollout: leader_index=%d out of range for n=%d", d->leader_index, d->n);
    if ((d->flags & HVP_ENV_REAL_VEHICLE_REF) && d->leader_index != 0)
        return fail(-4, "rollout: real_vehicle_as_reference needs leader_index 0 (env.py:60-63)");
    VehicleModel M;
    P.n = d->n; P.leader_index = d->leader_index; P.flags = d->flags;
    P.scen_per_block = ROLLOUT_BLOCK / d->n;
    P.d0 = d->d0; P.t0 = d->t0; P.d_safe = d->d_safe;
    memcpy(P.tr_t, M.tr_t, sizeof P.tr_t);
    memcpy(P.tr_v, M.tr_v, sizeof P.tr_v);
    for (int j = 0; j < 6; ++j) {
        P.inv_rise[j] = 1.0 / (M.tr_v[j][1] - M.tr_v[j][0]);
        P.inv_fall[j] = 1.0 / (M.tr_v[j][3] - M.tr_v[j][2]);
    }
    M.gear_limits(P.lim);
    P.c_fric = M.c_fric; P.mug = M.mu * M.grav; P.default_mass = 800.0;
    { int ex; P.fric_pow2 = (frexp(M.c_fric, &ex) == 0.5) ? 1 : 0; }
    P.inv_c_fric = 1.0 / M.c_fric;
    return 0;
}

extern "C" int hvp_rollout_step_dev(hvp_ctx* c, const hvp_env_desc* desc, int64_t batch, const double* x,
                                    const double* u, const int32_t* gear, const double* mass,
                                    const double* leader, double* x_out, double* cost, uint8_t* viol,
                                    int32_t* err, void* stream) {
    if (!c) return fail(-1, "rollout: ctx is NULL");
    if (batch < 0) return fail(-4, "rollout: negative batch");
    RolloutParams P;
    int rc = fill_rollout_params(P, desc);
    if (rc) return rc;
    if (batch == 0) return 0;
    if (!x || !u || !leader || !x_out || !cost || !viol || !err) return fail(-1, "rollout: NULL array argument");
    if (((uintptr_t)x & 15) || ((uintptr_t)x_out & 15)) return fail(-5, "rollout: x / x_out must be 16-byte aligned");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    CUDA_TRY(hvp_mark(c, c->ev0, st, false));
    CUDA_TRY(launch_rollout(P, batch, x, u, gear, mass, leader, x_out, cost, viol, err, st));
    CUDA_TRY(hvp_mark(c, c->ev1, st, true));
    c->launches += 1;
    return 0;
}

extern "C" int hvp_rollout_step_host(hvp_ctx* c, const hvp_env_desc* desc, int64_t batch, const double* x,
                                     const double* u, const int32_t* gear, const double* mass,
                                     const double* leader, double* x_out, double* cost, uint8_t* viol,
                                     int32_t* err) {
    if (!c) return fail(-1, "rollout: ctx is NULL");
    if (!desc) return fail(-1, "rollout: desc is NULL");
    if (batch < 0) return fail(-4, "rollout: negative batch");
    if (batch == 0) return 0;
    if (!x || !u || !leader || !x_out || !cost || !viol || !err) return fail(-1, "rollout: NULL array argument");
    {   // validate the descriptor BEFORE buffer sizes are derived from it (n <= 0 would wrap the size_t arithmetic)
        RolloutParams chk;
        const int rc0 = fill_rollout_params(chk, desc);
        if (rc0) return rc0;
    }
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t n = (size_t)desc->n, B = (size_t)batch;
    const bool mass_per = desc->flags & HVP_ENV_MASS_PER_SCENARIO;
    const size_t bx = align256(B * 2 * n * 8), bu = align256(B * n * 8), bg = gear ? align256(B * n * 4) : 0;
    const size_t bm = mass ? align256((mass_per ? B : 1) * n * 8) : 0, bl = align256(B * 16);
    const size_t bc = align256(B * 8), bv = align256(B), be = align256(B * 4);
    int rc = ensure_dbuf(c, bx * 2 + bu + bg + bm + bl + bc + bv + be);
    if (rc) return rc;
    char* q = c->dbuf;
    double* dx = (double*)q; q += bx;
    double* dxo = (double*)q; q += bx;
    double* du = (double*)q; q += bu;
    int32_t* dg = gear ? (int32_t*)q : nullptr; q += bg;
    double* dm = mass ? (double*)q : nullptr; q += bm;
    double* dl = (double*)q; q += bl;
    double* dc = (double*)q; q += bc;
    uint8_t* dv = (uint8_t*)q; q += bv;
    int32_t* de = (int32_t*)q;
    cudaStream_t st = c->stream;
    CUDA_TRY(cudaMemcpyAsync(dx, x, B * 2 * n * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(du, u, B * n * 8, cudaMemcpyHostToDevice, st));
    if (gear) CUDA_TRY(cudaMemcpyAsync(dg, gear, B * n * 4, cudaMemcpyHostToDevice, st));
    if (mass) CUDA_TRY(cudaMemcpyAsync(dm, mass, (mass_per ? B : 1) * n * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dl, leader, B * 16, cudaMemcpyHostToDevice, st));
    rc = hvp_rollout_step_dev(c, desc, batch, dx, du, dg, dm, dl, dxo, dc, dv, de, st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(x_out, dxo, B * 2 * n * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(cost, dc, B * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(viol, dv, B, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(err, de, B * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

// ------------------------------------------------------------------------------------------
// local MIQP
// ------------------------------------------------------------------------------------------
// counter and adoption scratch of a launch on `st` (see hvp_ctx): per stream, allocated on first use
static int launch_slot(hvp_ctx* c, cudaStream_t st, unsigned long long** counter, double** scratch, int* slot) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    const bool capturing = cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone;
    for (int i = 0; i < c->n_slots; ++i)
        if (c->slot_stream[i] == st) {
            c->slot_use[i] = ++c->use_clock;
            if (capturing) c->slot_pinned[i] = true;
            *counter = c->counters + i; *scratch = c->slot_scratch[i]; *slot = i;
            return 0;
        }
    if (!capturing) {                                       // no allocation, no event query while the stream is being captured
        int i = -1;
        if (c->n_slots < HVP_STREAM_SLOTS) {
            double* p = nullptr;
            CUDA_TRY(cudaMalloc(&p, HVP_STEAL_SLOT_DOUBLES * sizeof(double)));
            i = c->n_slots;
            if (cudaEventCreateWithFlags(&c->slot_done[i], cudaEventDisableTiming) != cudaSuccess) { cudaFree(p); return fail(-3, "launch_slot: cudaEventCreate failed"); }
            c->slot_scratch[i] = p; c->slot_pinned[i] = false;
            ++c->n_slots;
        } else {
            for (int k = 0; k < c->n_slots; ++k) {          // least recently used slot whose last launch has finished
                if (c->slot_pinned[k] || (i >= 0 && c->slot_use[k] >= c->slot_use[i])) continue;
                if (cudaEventQuery(c->slot_done[k]) == cudaSuccess) i = k;
            }
            cudaGetLastError();                             // cudaErrorNotReady of a busy slot is not an error
        }
        if (i >= 0) {
            c->slot_stream[i] = st; c->slot_use[i] = ++c->use_clock;
            *counter = c->counters + i; *scratch = c->slot_scratch[i]; *slot = i;
            return 0;
        }
    }
    *counter = c->counters + HVP_STREAM_SLOTS + (c->counter_next++ % HVP_COUNTER_RING);
    *scratch = nullptr; *slot = -1;
    return 0;
}

// after the launch that used `slot`: remember when it is done (not while capturing: the event would belong to the graph)
static void slot_mark(hvp_ctx* c, int slot, cudaStream_t st) {
    if (slot < 0) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) return;
    cudaEventRecord(c->slot_done[slot], st);
}

static int check_local_desc(const hvp_local_desc* d) {
    if (!d) return fail(-1, "local_miqp: desc is NULL");
    if (d->N < 2 || d->N > 12) return fail(-4, "local_miqp: N=%d out of range [2,12]", d->N);
    if (d->max_nodes < 0) return fail(-4, "local_miqp: negative max_nodes");
    if (!(d->mip_gap >= 0.0) || !(d->mip_gap < 1.0)) return fail(-4, "local_miqp: mip_gap must be in [0, 1)");
    if (!(d->time_limit_ms >= 0.0)) return fail(-4, "local_miqp: negative time_limit_ms");
    return 0;
}

extern "C" int hvp_local_miqp_dev(hvp_ctx* c, const hvp_local_desc* desc, int64_t batch, const int32_t* flags,
                                  const double* mass, const double* x0, const double* xf, const double* xb,
                                  const double* xl, double* u, double* x, int32_t* modes, double* obj,
                                  int32_t* status, int32_t* nodes, int32_t* qp_iters, void* stream) {
    if (!c) return fail(-1, "local_miqp: ctx is NULL");
    int rc = check_local_desc(desc);
    if (rc) return rc;
    if (batch < 0) return fail(-4, "local_miqp: negative batch");
    if (batch == 0) return 0;
    if (!flags || !mass || !x0 || !u || !x || !modes || !obj || !status || !nodes)
        return fail(-1, "local_miqp: NULL array argument");
    LocalParams P;
    fill_local_params(P, desc->N, desc->d0, desc->t0, desc->tight, desc->max_nodes, desc->mip_gap, desc->time_limit_ms);
    P.hint = desc->modes_hint;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    CUDA_TRY(hvp_mark(c, c->ev0, st, false));
    unsigned long long* counter = nullptr;
    double* steal = nullptr;
    int slot = -1;
    rc = launch_slot(c, st, &counter, &steal, &slot);
    if (rc) return rc;
    CUDA_TRY(launch_local_miqp(P, counter, steal, batch, flags, mass, x0, xf, xb, xl, u, x, modes, obj, status, nodes, qp_iters, st));
    slot_mark(c, slot, st);
    CUDA_TRY(hvp_mark(c, c->ev1, st, true));
    c->launches += 1;
    return 0;
}

extern "C" int hvp_local_miqp_host(hvp_ctx* c, const hvp_local_desc* desc, int64_t batch, const int32_t* flags,
                                   const double* mass, const double* x0, const double* xf, const double* xb,
                                   const double* xl, double* u, double* x, int32_t* modes, double* obj,
                                   int32_t* status, int32_t* nodes, int32_t* qp_iters) {
    if (!c) return fail(-1, "local_miqp: ctx is NULL");
    int rc = check_local_desc(desc);
    if (rc) return rc;
    if (batch < 0) return fail(-4, "local_miqp: negative batch");
    if (batch == 0) return 0;
    if (!flags || !mass || !x0 || !u || !x || !modes || !obj || !status || !nodes)
        return fail(-1, "local_miqp: NULL array argument");
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t B = (size_t)batch, N = (size_t)desc->N, S = 2 * (N + 1);
    const size_t bf = align256(B * 4), bm = align256(B * 8), b0 = align256(B * 16), br = align256(B * S * 8);
    const size_t bu = align256(B * N * 8), bmo = align256(B * N * 4), bo = align256(B * 8), bs = align256(B * 4);
    const size_t nref = (xf ? 1 : 0) + (xb ? 1 : 0) + (xl ? 1 : 0);
    rc = ensure_dbuf(c, bf + bm + b0 + br * nref + bu + br + bmo + bo + bs * 3);
    if (rc) return rc;
    char* q = c->dbuf;
    int32_t* dfl = (int32_t*)q; q += bf;
    double* dma = (double*)q; q += bm;
    double* dx0 = (double*)q; q += b0;
    double* dxf = xf ? (double*)q : nullptr; q += xf ? br : 0;
    double* dxb = xb ? (double*)q : nullptr; q += xb ? br : 0;
    double* dxl = xl ? (double*)q : nullptr; q += xl ? br : 0;
    double* du = (double*)q; q += bu;
    double* dx = (double*)q; q += br;
    int32_t* dmo = (int32_t*)q; q += bmo;
    double* dob = (double*)q; q += bo;
    int32_t* dst = (int32_t*)q; q += bs;
    int32_t* dno = (int32_t*)q; q += bs;
    int32_t* dit = qp_iters ? (int32_t*)q : nullptr;
    cudaStream_t st = c->stream;
    const size_t in_bytes = (size_t)((char*)du - c->dbuf), all_bytes = (size_t)(q - c->dbuf) + bs;
    if (all_bytes <= (256u << 10)) {
        // latency path (one scenario-timestep = a handful of MIQPs): stage through a pinned mirror of the
        // device buffer so the call costs ONE H2D and ONE D2H copy instead of 6 + 7 pageable ones
        if (c->hcap < all_bytes) {
            if (c->hbuf) { CUDA_TRY(cudaFreeHost(c->hbuf)); c->hbuf = nullptr; c->hcap = 0; }
            CUDA_TRY(cudaMallocHost(&c->hbuf, 256u << 10));
            c->hcap = 256u << 10;
        }
        char* h = c->hbuf;
        auto off = [&](const void* d) { return (size_t)((const char*)d - c->dbuf); };
        memcpy(h + off(dfl), flags, B * 4); memcpy(h + off(dma), mass, B * 8); memcpy(h + off(dx0), x0, B * 16);
        if (xf) memcpy(h + off(dxf), xf, B * S * 8);
        if (xb) memcpy(h + off(dxb), xb, B * S * 8);
        if (xl) memcpy(h + off(dxl), xl, B * S * 8);
        CUDA_TRY(cudaMemcpyAsync(c->dbuf, h, in_bytes, cudaMemcpyHostToDevice, st));
        rc = hvp_local_miqp_dev(c, desc, batch, dfl, dma, dx0, dxf, dxb, dxl, du, dx, dmo, dob, dst, dno, dit, st);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(h + in_bytes, c->dbuf + in_bytes, all_bytes - in_bytes, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        memcpy(u, h + off(du), B * N * 8); memcpy(x, h + off(dx), B * S * 8); memcpy(modes, h + off(dmo), B * N * 4);
        memcpy(obj, h + off(dob), B * 8); memcpy(status, h + off(dst), B * 4); memcpy(nodes, h + off(dno), B * 4);
        if (qp_iters) memcpy(qp_iters, h + off(dit), B * 4);
        return 0;
    }
    // Large batches go through in CHUNKS on three side streams: the H2D copy of one chunk, the kernel of another and
    // the D2H copy of a third overlap (PCIe is full duplex), and so do the kernels themselves at their ends -- the
    // persistent kernel of a chunk drains its last, longest trees while the next chunk's CTAs take the freed slots.
    // (Host buffers should be pinned for the copies to be asynchronous; pageable memory still works, serialised.)
    static const size_t CH = getenv("HVP_HOST_CHUNK") ? (size_t)atoll(getenv("HVP_HOST_CHUNK")) : 98304;
    if (B >= 2 * CH) {
        if (!c->side_ok) {
            for (int i = 0; i < 3; ++i) CUDA_TRY(cudaStreamCreateWithFlags(&c->side[i], cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&c->side_ev, cudaEventDisableTiming));
            c->side_ok = true;
        }
        CUDA_TRY(cudaEventRecord(c->side_ev, st));              // the side streams start after earlier work on `st`
        for (int i = 0; i < 3; ++i) CUDA_TRY(cudaStreamWaitEvent(c->side[i], c->side_ev, 0));
        CUDA_TRY(cudaEventRecord(c->ev0, st));
        int k = 0;
        for (size_t o = 0; o < B; o += CH, ++k) {
            const size_t nb = (B - o < CH + CH / 2) ? B - o : CH;   // the last chunk takes a short remainder with it
            cudaStream_t ss = c->side[k % 3];
            CUDA_TRY(cudaMemcpyAsync(dfl + o, flags + o, nb * 4, cudaMemcpyHostToDevice, ss));
            CUDA_TRY(cudaMemcpyAsync(dma + o, mass + o, nb * 8, cudaMemcpyHostToDevice, ss));
            CUDA_TRY(cudaMemcpyAsync(dx0 + 2 * o, x0 + 2 * o, nb * 16, cudaMemcpyHostToDevice, ss));
            if (xf) CUDA_TRY(cudaMemcpyAsync(dxf + S * o, xf + S * o, nb * S * 8, cudaMemcpyHostToDevice, ss));
            if (xb) CUDA_TRY(cudaMemcpyAsync(dxb + S * o, xb + S * o, nb * S * 8, cudaMemcpyHostToDevice, ss));
            if (xl) CUDA_TRY(cudaMemcpyAsync(dxl + S * o, xl + S * o, nb * S * 8, cudaMemcpyHostToDevice, ss));
            LocalParams P;
            fill_local_params(P, desc->N, desc->d0, desc->t0, desc->tight, desc->max_nodes, desc->mip_gap, desc->time_limit_ms);
            unsigned long long* counter = nullptr;
            double* steal = nullptr;
            int slot = -1;
            rc = launch_slot(c, ss, &counter, &steal, &slot);
            if (rc) return rc;
            CUDA_TRY(launch_local_miqp(P, counter, steal, (int64_t)nb, dfl + o, dma + o, dx0 + 2 * o, dxf ? dxf + S * o : nullptr,
                                       dxb ? dxb + S * o : nullptr, dxl ? dxl + S * o : nullptr, du + N * o, dx + S * o,
                                       dmo + N * o, dob + o, dst + o, dno + o, dit ? dit + o : nullptr, ss));
            slot_mark(c, slot, ss);
            c->launches += 1;
            CUDA_TRY(cudaMemcpyAsync(u + N * o, du + N * o, nb * N * 8, cudaMemcpyDeviceToHost, ss));
            CUDA_TRY(cudaMemcpyAsync(x + S * o, dx + S * o, nb * S * 8, cudaMemcpyDeviceToHost, ss));
            CUDA_TRY(cudaMemcpyAsync(modes + N * o, dmo + N * o, nb * N * 4, cudaMemcpyDeviceToHost, ss));
            CUDA_TRY(cudaMemcpyAsync(obj + o, dob + o, nb * 8, cudaMemcpyDeviceToHost, ss));
            CUDA_TRY(cudaMemcpyAsync(status + o, dst + o, nb * 4, cudaMemcpyDeviceToHost, ss));
            CUDA_TRY(cudaMemcpyAsync(nodes + o, dno + o, nb * 4, cudaMemcpyDeviceToHost, ss));
            if (qp_iters) CUDA_TRY(cudaMemcpyAsync(qp_iters + o, dit + o, nb * 4, cudaMemcpyDeviceToHost, ss));
            if (nb != CH) break;
        }
        for (int i = 0; i < 3; ++i) {                            // `st` continues after all three side streams
            CUDA_TRY(cudaEventRecord(c->side_ev, c->side[i]));
            CUDA_TRY(cudaStreamWaitEvent(st, c->side_ev, 0));
        }
        CUDA_TRY(hvp_mark(c, c->ev1, st, true));
        CUDA_TRY(cudaStreamSynchronize(st));
        return 0;
    }
    CUDA_TRY(cudaMemcpyAsync(dfl, flags, B * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dma, mass, B * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dx0, x0, B * 16, cudaMemcpyHostToDevice, st));
    if (xf) CUDA_TRY(cudaMemcpyAsync(dxf, xf, B * S * 8, cudaMemcpyHostToDevice, st));
    if (xb) CUDA_TRY(cudaMemcpyAsync(dxb, xb, B * S * 8, cudaMemcpyHostToDevice, st));
    if (xl) CUDA_TRY(cudaMemcpyAsync(dxl, xl, B * S * 8, cudaMemcpyHostToDevice, st));
    rc = hvp_local_miqp_dev(c, desc, batch, dfl, dma, dx0, dxf, dxb, dxl, du, dx, dmo, dob, dst, dno, dit, st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(u, du, B * N * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(x, dx, B * S * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(modes, dmo, B * N * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(obj, dob, B * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(status, dst, B * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(nodes, dno, B * 4, cudaMemcpyDeviceToHost, st));
    if (qp_iters) CUDA_TRY(cudaMemcpyAsync(qp_iters, dit, B * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

// ------------------------------------------------------------------------------------------
// measurement helper: FP64 FMA peak
// ------------------------------------------------------------------------------------------
extern "C" int hvp_microbench_fp64(hvp_ctx* c, int iters, double* tflops) {
    if (!c || !tflops) return fail(-1, "microbench: NULL argument");
    if (iters < 1) return fail(-4, "microbench: iters < 1");
    CUDA_TRY(cudaSetDevice(c->device));
    double* sink = nullptr;
    CUDA_TRY(cudaMalloc(&sink, 8 * 1024));
    int blocks = 0, threads = 0;
    cudaStream_t st = c->stream;
    CUDA_TRY(launch_fp64_microbench(iters, sink, &blocks, &threads, st));   // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(c->ev0, st));
        CUDA_TRY(launch_fp64_microbench(iters, sink, &blocks, &threads, st));
        CUDA_TRY(cudaEventRecord(c->ev1, st));
        CUDA_TRY(cudaEventSynchronize(c->ev1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (ms < best) best = ms;
        c->launches += 1;
    }
    c->launches += 1;
    CUDA_TRY(cudaFree(sink));
    const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    return 0;
}

// measurement helper: shared-memory read bandwidth (GB/s over the whole device)
extern "C" int hvp_microbench_smem(hvp_ctx* c, int iters, double* gbs) {
    if (!c || !gbs) return fail(-1, "microbench: NULL argument");
    if (iters < 1) return fail(-4, "microbench: iters < 1");
    CUDA_TRY(cudaSetDevice(c->device));
    double* sink = nullptr;
    CUDA_TRY(cudaMalloc(&sink, 8 * 1024));
    int blocks = 0, threads = 0;
    cudaStream_t st = c->stream;
    CUDA_TRY(launch_smem_microbench(iters, sink, &blocks, &threads, st));   // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(c->ev0, st));
        CUDA_TRY(launch_smem_microbench(iters, sink, &blocks, &threads, st));
        CUDA_TRY(cudaEventRecord(c->ev1, st));
        CUDA_TRY(cudaEventSynchronize(c->ev1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (ms < best) best = ms;
    }
    c->launches += 6;
    CUDA_TRY(cudaFree(sink));
    const double bytes = 4.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
    *gbs = bytes / (best * 1e-3) / 1e9;
    return 0;
}
