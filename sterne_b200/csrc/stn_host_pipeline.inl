// Host side, part 2: the host-buffer pipeline (pinned / pageable staging, chunk schedule, three slots).
// Included by sterne_b200.cu.
namespace {

// ---- host-buffer pipeline ---------------------------------------------------------------
// true when the CUDA runtime knows `p` as pinned (page-locked) host memory, i.e. a DMA target
bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

int host_pipe_init(StnCtx* ctx) {
    for (int i = 0; i < StnCtx::kSlots; ++i) {
        if (!ctx->streams[i]) STN_CUDA(cudaStreamCreateWithFlags(&ctx->streams[i], cudaStreamNonBlocking));
        if (!ctx->h2d_done[i]) STN_CUDA(cudaEventCreateWithFlags(&ctx->h2d_done[i], cudaEventDisableTiming));
        if (!ctx->d2h_done[i]) STN_CUDA(cudaEventCreateWithFlags(&ctx->d2h_done[i], cudaEventDisableTiming));
    }
    return STN_OK;
}

void host_pipe_drain(StnCtx* ctx) {
    for (int i = 0; i < StnCtx::kSlots; ++i)
        if (ctx->streams[i]) cudaStreamSynchronize(ctx->streams[i]);
}

// Rows of the k-th chunk: a short first chunk (its H2D copy is the only one nothing overlaps),
// growing fourfold up to 2^20 rows (every chunk boundary costs a partly filled wave of CTAs).
#ifndef STN_HOST_CHUNK_LOG2
#define STN_HOST_CHUNK_LOG2 20
#endif
constexpr int64_t kChunkRowsMax = (int64_t)1 << STN_HOST_CHUNK_LOG2;

// memcpy of a large block on a few threads: one core moves ~12 GB/s, which is below what the
// kernel consumes for short-series problems and would make pageable buffers the bottleneck
void big_memcpy(void* dst, const void* src, size_t bytes) {
    constexpr size_t kPerThread = (size_t)8 << 20;
    int nt = (int)(bytes / kPerThread);
    if (nt > 4) nt = 4;
    if (nt <= 1) {
        std::memcpy(dst, src, bytes);
        return;
    }
    const size_t part = ((bytes / nt) + 63) & ~(size_t)63;
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) {
        const size_t lo = (size_t)t * part;
        const size_t len = lo >= bytes ? 0 : (t == nt - 1 ? bytes - lo : part);
        if (len) pool.emplace_back([=]() { std::memcpy((char*)dst + lo, (const char*)src + lo, len); });
    }
    std::memcpy(dst, src, part < bytes ? part : bytes);
    for (std::thread& th : pool) th.join();
}
// Tuning overrides (environment, read once): STN_HOST_FIRST_LOG2 = log2 of the first chunk's rows,
// STN_HOST_GROWTH = factor between consecutive chunks.
struct HostSchedule {
    int64_t first = 0;
    int growth = 0;
    HostSchedule() {
        if (const char* e = std::getenv("STN_HOST_FIRST_LOG2")) first = (int64_t)1 << std::atoi(e);
        if (const char* e = std::getenv("STN_HOST_GROWTH")) growth = std::atoi(e);
    }
};
const HostSchedule& host_schedule() {
    static const HostSchedule s;
    return s;
}

int64_t first_chunk_rows(int64_t n) {
    const HostSchedule& hs = host_schedule();
    if (hs.first > 0) return hs.first < n ? hs.first : n;
    int64_t r = n / 32;
    if (r > ((int64_t)1 << 15)) r = (int64_t)1 << 15;
    if (r < 4096) r = 4096;
    return r < n ? r : n;
}

// Next chunk: fourfold growth when this device has the host's PCIe bandwidth to itself (~50 GB/s: the copy of
// a 4x larger chunk still hides under the kernel of the current one), twofold when several devices copy at
// once (stn_multi_*, or stn_host_hint from a multi-process launcher).  Chunks of a wave or more are cut at
// whole waves of the big kernel so that only the last chunk ends in a partly filled wave.
int64_t next_chunk_rows(const StnCtx* ctx, int64_t rows) {
    // (twofold growth for pageable, staged input was tried as well: 55.1 ms instead of 52.6 at 2^22 rows)
    int g = host_schedule().growth > 1 ? host_schedule().growth : (ctx->host_share > 1 ? 2 : 4);
    int64_t r = rows >= kChunkRowsMax ? kChunkRowsMax : (rows * g < kChunkRowsMax ? rows * g : kChunkRowsMax);
    const int bigR = ctx->efac_heavy ? STN_RBIG_EFAC : STN_RBIG;
    const int64_t wave = (int64_t)ctx->sm_count * minblocks_for(bigR) * threads_for(bigR) * bigR;
    if (r >= wave) r = (r / wave) * wave;
    return r;
}

// One slice of a host batch on this context's device: H2D, kernel(s), D2H, synchronised on return.
// The caller has made ctx->device current and resolved `plan` from the WHOLE batch.
int run_host(StnCtx* ctx, const double* theta_host, int64_t n, int64_t ld, int layout, int mode, int plan,
             double* out_host) {
    const int P = ctx->prob.n_params;
    int rc = host_pipe_init(ctx);
    if (rc) return rc;
    // Small batches (a sampler evaluating one point, or a few walkers) are latency-bound: skip both
    // cudaMemcpy calls by letting the kernel read the parameters from, and write log L to, pinned
    // mapped host memory.  One launch + one stream synchronise.
    // (limits tunable through the environment; measured with tools/latency_host.py: raising them to 8192 rows / 2 MB
    // does not help 6-parameter batches and doubles the latency of 16-parameter ones -- 82 -> 152 us at 4096 rows)
    static const int64_t kSmallBytes = std::getenv("STN_SMALL_BYTES") ? std::atoll(std::getenv("STN_SMALL_BYTES")) : 64 * 1024;
    static const int64_t kSmallRows = std::getenv("STN_SMALL_ROWS") ? std::atoll(std::getenv("STN_SMALL_ROWS")) : 2048;
    if (layout == STN_LAYOUT_AOS && n <= kSmallRows && n * P * (int64_t)sizeof(double) <= kSmallBytes) {
        if (!ctx->small_in) STN_CUDA(cudaHostAlloc((void**)&ctx->small_in, kSmallBytes, cudaHostAllocMapped));
        if (!ctx->small_out) STN_CUDA(cudaHostAlloc((void**)&ctx->small_out, kSmallRows * sizeof(double), cudaHostAllocMapped));
        for (int64_t i = 0; i < n; ++i) std::memcpy(ctx->small_in + i * P, theta_host + i * ld, (size_t)P * sizeof(double));
        rc = launch_loglike(ctx, ctx->small_in, n, P, layout, mode, plan, ctx->small_out, ctx->streams[0], 0, 1);
        const cudaError_t e = cudaStreamSynchronize(ctx->streams[0]);
        if (rc) return rc;
        if (e != cudaSuccess) return fail(STN_ERR_CUDA, "small-batch launch failed: %s", cudaGetErrorString(e));
        std::memcpy(out_host, ctx->small_out, (size_t)n * sizeof(double));
        return STN_OK;
    }
    const bool pin_in = is_pinned(theta_host);
    const bool pin_out = is_pinned(out_host);
    const int64_t cap = n < kChunkRowsMax ? n : kChunkRowsMax;
    // device staging (and, for pageable caller buffers, the library's own pinned ring)
    if (ctx->stage_rows < cap) {
        host_pipe_drain(ctx);
        for (int i = 0; i < StnCtx::kSlots; ++i) {
            if (ctx->stage_theta[i]) cudaFree(ctx->stage_theta[i]);
            if (ctx->stage_out[i]) cudaFree(ctx->stage_out[i]);
            ctx->stage_theta[i] = ctx->stage_out[i] = nullptr;
        }
        ctx->stage_rows = 0;
        for (int i = 0; i < StnCtx::kSlots; ++i) {
            STN_CUDA(cudaMalloc(&ctx->stage_theta[i], (size_t)cap * P * sizeof(double)));
            STN_CUDA(cudaMalloc(&ctx->stage_out[i], (size_t)cap * sizeof(double)));
        }
        ctx->stage_rows = cap;
    }
    if ((!pin_in || !pin_out) && ctx->pin_rows < cap) {
        host_pipe_drain(ctx);
        for (int i = 0; i < StnCtx::kSlots; ++i) {
            if (ctx->pin_theta[i]) cudaFreeHost(ctx->pin_theta[i]);
            if (ctx->pin_out[i]) cudaFreeHost(ctx->pin_out[i]);
            ctx->pin_theta[i] = ctx->pin_out[i] = nullptr;
        }
        ctx->pin_rows = 0;
        for (int i = 0; i < StnCtx::kSlots; ++i) {
            STN_CUDA(cudaHostAlloc((void**)&ctx->pin_theta[i], (size_t)cap * P * sizeof(double), cudaHostAllocDefault));
            STN_CUDA(cudaHostAlloc((void**)&ctx->pin_out[i], (size_t)cap * sizeof(double), cudaHostAllocDefault));
        }
        ctx->pin_rows = cap;
    }

    struct Pending { int64_t lo, m; bool live; } pend[StnCtx::kSlots] = {};
    auto body = [&]() -> int {
        int64_t rows = first_chunk_rows(n);
        int k = 0;
        for (int64_t lo = 0; lo < n; ++k) {
            const int64_t m = (n - lo) < rows ? (n - lo) : rows;
            const int slot = k % StnCtx::kSlots;
            cudaStream_t st = ctx->streams[slot];
            double* dth = ctx->stage_theta[slot];
            double* dout = ctx->stage_out[slot];
            // pageable output: the chunk that used this slot last is copied out before the slot is reused
            if (pend[slot].live) {
                STN_CUDA(cudaEventSynchronize(ctx->d2h_done[slot]));
                std::memcpy(out_host + pend[slot].lo, ctx->pin_out[slot], (size_t)pend[slot].m * sizeof(double));
                pend[slot].live = false;
            }
            const double* src;
            int64_t src_ld;                 // leading dimension of the host image that is DMA'd
            if (pin_in) {
                src = layout == STN_LAYOUT_AOS ? theta_host + lo * ld : theta_host + lo;
                src_ld = ld;
            } else {
                // pageable input: compact copy into this slot's pinned buffer (the previous DMA out of it is done
                // once h2d_done has fired), then DMA from there while the next chunk is being copied
                if (k >= StnCtx::kSlots) STN_CUDA(cudaEventSynchronize(ctx->h2d_done[slot]));
                double* hp = ctx->pin_theta[slot];
                if (layout == STN_LAYOUT_AOS) {
                    if (ld == P) big_memcpy(hp, theta_host + lo * ld, (size_t)m * P * sizeof(double));
                    else for (int64_t i = 0; i < m; ++i) std::memcpy(hp + i * P, theta_host + (lo + i) * ld, (size_t)P * sizeof(double));
                    src_ld = P;
                } else {
                    for (int c = 0; c < P; ++c) big_memcpy(hp + (int64_t)c * m, theta_host + (int64_t)c * ld + lo, (size_t)m * sizeof(double));
                    src_ld = m;
                }
                src = hp;
            }
            int64_t dld;
            if (layout == STN_LAYOUT_AOS && src_ld == P) {
                STN_CUDA(cudaMemcpyAsync(dth, src, (size_t)m * P * sizeof(double), cudaMemcpyHostToDevice, st));
                dld = P;
            } else if (layout == STN_LAYOUT_AOS) {
                // padded rows: keep only the first P columns of each
                STN_CUDA(cudaMemcpy2DAsync(dth, (size_t)P * sizeof(double), src, (size_t)src_ld * sizeof(double),
                                           (size_t)P * sizeof(double), (size_t)m, cudaMemcpyHostToDevice, st));
                dld = P;
            } else {
                STN_CUDA(cudaMemcpy2DAsync(dth, (size_t)m * sizeof(double), src, (size_t)src_ld * sizeof(double),
                                           (size_t)m * sizeof(double), (size_t)P, cudaMemcpyHostToDevice, st));
                dld = m;
            }
            if (!pin_in) STN_CUDA(cudaEventRecord(ctx->h2d_done[slot], st));
            const int r = launch_loglike(ctx, dth, m, dld, layout, mode, plan, dout, st, 0, 1 + slot);
            if (r) return r;
            if (pin_out) {
                STN_CUDA(cudaMemcpyAsync(out_host + lo, dout, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
            } else {
                STN_CUDA(cudaMemcpyAsync(ctx->pin_out[slot], dout, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
                STN_CUDA(cudaEventRecord(ctx->d2h_done[slot], st));
                pend[slot].lo = lo; pend[slot].m = m; pend[slot].live = true;
            }
            lo += m;
            rows = next_chunk_rows(ctx, rows);
        }
        // the chunks still in flight, oldest first
        for (int j = 0; j < StnCtx::kSlots; ++j) {
            const int slot = (k + j) % StnCtx::kSlots;
            if (pend[slot].live) {
                STN_CUDA(cudaEventSynchronize(ctx->d2h_done[slot]));
                std::memcpy(out_host + pend[slot].lo, ctx->pin_out[slot], (size_t)pend[slot].m * sizeof(double));
                pend[slot].live = false;
            }
        }
        for (int i = 0; i < StnCtx::kSlots; ++i) STN_CUDA(cudaStreamSynchronize(ctx->streams[i]));
        return STN_OK;
    };
    rc = body();
    if (rc) host_pipe_drain(ctx);     // nothing of this call may still touch the caller's buffers after an error
    return rc;
}

}  // namespace
