// Host side, part 1: error reporting, the context structs, uploads, execution plans and kernel launches.
// Included by sterne_b200.cu.
namespace {

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define STN_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(STN_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

}  // namespace

struct StnCtx {
    int device = 0;
    int sm_count = 0;
    ProbDev prob{};
    std::vector<void*> dev_allocs;
    std::vector<int> n_epochs;          // per source
    int max_epochs = 0;
    ProbDev prior_prob{};               // copy of prob with the prior switched on (stn_set_prior)
    bool has_prior = false;
    std::vector<void*> prior_allocs;
    bool efac_heavy = false;            // at least half of the epochs belong to EFAC sources
    // zero-copy staging for small host batches (pinned, mapped: the kernel reads / writes it directly)
    double* small_in = nullptr;
    double* small_out = nullptr;
    static constexpr int kSlots = 3;    // host pipeline depth: H2D of chunk i+1, kernel of chunk i, D2H of chunk i-1
    double* scratch[1 + kSlots] = {};   // split partial sums: [0] device API, [1 + slot] host pipeline
    int64_t scratch_elems[1 + kSlots] = {};
    int64_t launches = 0;
    // *_host pipeline: one stream, one device staging pair and (for pageable caller buffers) one pinned
    // staging pair per slot
    cudaStream_t streams[kSlots] = {};
    cudaEvent_t h2d_done[kSlots] = {};
    cudaEvent_t d2h_done[kSlots] = {};
    double* stage_theta[kSlots] = {};
    double* stage_out[kSlots] = {};
    int64_t stage_rows = 0;
    double* pin_theta[kSlots] = {};
    double* pin_out[kSlots] = {};
    int64_t pin_rows = 0;
    int host_share = 1;                 // devices DMA-ing from this host at the same time (stn_host_hint)
    cudaStream_t timing_stream = nullptr;
    double* gather_buf = nullptr;       // stn_multi_loglike_dev: this device's slice before the peer copy
    int64_t gather_elems = 0;
    double* sampler_buf = nullptr;      // stn_ensemble_run (three-launch engine): proposals, stretch factors, log-posteriors
    int64_t sampler_elems = 0;
    cudaEvent_t sampler_ev[2] = {};     // stn_multi_ensemble_run: "this device finished half-step h" (ping-pong)
    unsigned long long* sampler_nacc = nullptr;
};

// Several contexts of the same problem, one per device, driven as one (stn_multi_*).
struct StnMulti {
    std::vector<StnCtx*> ctx;
    std::vector<int> peer_ok;           // [a * G + b]: kernels on ctx[a]'s device may store to ctx[b]'s device
};

namespace {

template <typename T>
int upload(StnCtx* ctx, const T* host, size_t count, const T** dev_out) {
    void* d = nullptr;
    const size_t bytes = (count ? count : 1) * sizeof(T);
    STN_CUDA(cudaMalloc(&d, bytes));
    ctx->dev_allocs.push_back(d);
    if (count) STN_CUDA(cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *dev_out = static_cast<const T*>(d);
    return STN_OK;
}

#ifdef STN_DEBUG
// the `bytes` starting at device pointer p must lie inside one allocation known to the driver
int dbg_range(const void* p, size_t bytes, const char* what) {
    if (bytes == 0) return STN_OK;
    CUdeviceptr base = 0;
    size_t size = 0;
    const CUresult r = cuMemGetAddressRange(&base, &size, (CUdeviceptr)(uintptr_t)p);
    if (r != CUDA_SUCCESS) return fail(STN_ERR_INVALID, "STN_DEBUG: %s (%p) is not inside a CUDA allocation", what, p);
    const uintptr_t lo = (uintptr_t)p, hi = lo + bytes;
    if (lo < (uintptr_t)base || hi > (uintptr_t)base + size)
        return fail(STN_ERR_INVALID, "STN_DEBUG: %s needs %zu bytes at %p but its allocation is [%p, +%zu)", what, bytes,
                    p, (void*)(uintptr_t)base, size);
    return STN_OK;
}
size_t dbg_theta_bytes(int64_t n, int64_t ld, int layout, int P) {
    if (n == 0) return 0;
    return sizeof(double) * (size_t)(layout == STN_LAYOUT_AOS ? (n - 1) * ld + P : (int64_t)(P - 1) * ld + n);
}
#define STN_DBG_RANGE(p, bytes, what) do { const int r_ = dbg_range(p, bytes, what); if (r_) return r_; } while (0)
#else
#define STN_DBG_RANGE(p, bytes, what) ((void)0)
#endif

// ---- execution plan -------------------------------------------------------------------
// plan = lanes | (splits << 8).  lanes: lanes of a warp sharing one sample (epochs of a chunk
// interleaved over them); splits: CTAs sharing one tile of samples (the chunk list cut into
// contiguous ranges).  Both change the summation order, i.e. result bits, which is why the
// plan is explicit in the ABI; samples per thread (R) never does.
inline int plan_lanes(int plan) { return plan & 0xff; }
inline int plan_splits(int plan) { return (plan >> 8) & 0xffff; }
inline int make_plan(int lanes, int splits) { return lanes | (splits << 8); }

int big_R(const StnCtx* ctx) { return ctx->efac_heavy ? STN_RBIG_EFAC : STN_RBIG; }

// R used for a batch of n with `splits` CTAs per tile: the big R once its tiles fill most of a wave
int choose_R(const StnCtx* ctx, int64_t n, int lanes, int splits) {
    if (lanes != 1) return 1;
    const int R = big_R(ctx);
    const int64_t slots = (int64_t)ctx->sm_count * minblocks_for(R);
    const int64_t tile = (int64_t)threads_for(R) * R;
    const int64_t tiles = (n + tile - 1) / tile;
    // measured, variant C, plan lanes=1|splits=1: at 0.58 of a wave R = 1 wins (1.05 vs 1.37 ms), from 0.82 of a
    // wave on the big kernel does (1.36 vs 1.54 ms; 1.15 waves: 1.96 vs 2.03; 1.75 waves: 2.81 vs 3.01)
    return 10 * tiles * splits >= STN_BIGR_MIN_TENTH_WAVES * slots ? R : 1;
}

int pick_plan(const StnCtx* ctx, int64_t n) {
    if (n <= 0) return make_plan(1, 1);
    const int T = ctx->prob.n_chunks;
    const int R = big_R(ctx);
    const int64_t slots_big = (int64_t)ctx->sm_count * minblocks_for(R);
    const int64_t tile_big = (int64_t)threads_for(R) * R;
    const int64_t tiles_big = (n + tile_big - 1) / tile_big;
    if (tiles_big >= 32 * slots_big) return make_plan(1, 1);         // >= 32 waves: the last one is < 3 %
    // choose the number of splits that wastes the least of the last wave, given the tile count
    auto eff = [&](int64_t tiles, int64_t slots, int K) {
        const int64_t ctas = tiles * K;
        const int64_t waves = (ctas + slots - 1) / slots;
        const double fill = (double)ctas / (double)(waves * slots);
        const int per = (T + K - 1) / K;                              // chunks of the longest split
        const double balance = (double)T / (double)((int64_t)per * K);
        return fill * balance;
    };
    if (tiles_big * (int64_t)T >= 2 * slots_big) {                    // the big-R kernel can fill the GPU
        int bestK = 1;
        double best = eff(tiles_big, slots_big, 1);
        for (int K = 2; K <= T && K <= 64; ++K) {
            if (tiles_big * K < 2 * slots_big && K < T) continue;
            const double e = eff(tiles_big, slots_big, K);
            if (e > best * 1.02) { best = e; bestK = K; }
            if (tiles_big * K >= 32 * slots_big) break;
        }
        if (tiles_big * bestK >= 2 * slots_big) return make_plan(1, bestK);
    }
    // small batches: one sample per thread, as many splits as there are chunks allow, then lanes
    const int64_t slots = (int64_t)ctx->sm_count * 2;
    const int64_t tiles = (n + kThreads - 1) / kThreads;
    int K = 1;
    while (K < T && K < 64 && tiles * K < 2 * slots) ++K;
    int cap = 1;                                                       // lanes never outnumber epochs/8
    while (cap < 32 && cap * 8 <= ctx->max_epochs && cap * 8 <= kChunk) cap *= 2;
    int L = 1;
    while (L < cap && ((n * L + kThreads - 1) / kThreads) * K < 2 * slots) L *= 2;
    return make_plan(L, K);
}

template <int R, int L, int TPB = threads_for(R)>
int launch_fast(StnCtx* ctx, const ProbDev& prob, const double* theta, int64_t n, int64_t ld, int layout,
                double* out, cudaStream_t st, int splits, double* partial) {
    constexpr int per_cta = (TPB / L) * R;
    const int64_t grid = (n + per_cta - 1) / per_cta;
    if (grid > 0x7fffffffLL) return fail(STN_ERR_INVALID, "batch too large for one launch");
    dim3 g((unsigned)grid, (unsigned)splits, 1);
    (void)cudaGetLastError();      // other CUDA users of this process (torch) may have left a stale error
#if STN_DYNAMIC_STAGE
    static bool attr_set = false;                    // per instantiation
    if (!attr_set) {
        STN_CUDA(cudaFuncSetAttribute(stn_loglike_fast_kernel<R, L, TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kStageSmem));
        attr_set = true;
    }
#endif
    stn_loglike_fast_kernel<R, L, TPB><<<g, TPB, kStageSmem, st>>>(prob, theta, n, ld, layout, out, partial);
    STN_CUDA(cudaGetLastError());
    ctx->launches++;
    if (splits > 1) {
        const int64_t rg = (n + kThreads - 1) / kThreads;
        stn_reduce_splits_kernel<<<(unsigned)rg, kThreads, 0, st>>>(prob, theta, n, ld, layout, splits, partial, out);
        STN_CUDA(cudaGetLastError());
        ctx->launches++;
    }
    return STN_OK;
}

// scratch_slot: 0 for the device-buffer API, 1 + slot for the streams of the host pipeline
int launch_loglike(StnCtx* ctx, const double* theta, int64_t n, int64_t ld, int layout, int mode,
                   int plan, double* out, cudaStream_t st, int what = 0, int scratch_slot = 0,
                   double* const* extra_out = nullptr, int n_extra_out = 0) {
    // what: 0 = log L, 1 = log prior + log L, 2 = chi-square
    if (n == 0) return STN_OK;
    ProbDev prob = what == 1 ? ctx->prior_prob : ctx->prob;
    prob.out_chisq = what == 2 ? 1 : 0;
    prob.n_extra_out = n_extra_out;
    for (int k = 0; k < n_extra_out; ++k) prob.extra_out[k] = extra_out[k];
    if (mode == STN_MODE_STRICT) {
        STN_DBG_RANGE(theta, dbg_theta_bytes(n, ld, layout, ctx->prob.n_params), "theta");
        STN_DBG_RANGE(out, (size_t)n * sizeof(double), "out");
        const int64_t grid = (n + kThreads - 1) / kThreads;
        (void)cudaGetLastError();
        stn_loglike_strict_kernel<<<(unsigned)grid, kThreads, 0, st>>>(prob, theta, n, ld, layout, out);
        STN_CUDA(cudaGetLastError());
        ctx->launches++;
        return STN_OK;
    }
    if (plan == 0) plan = pick_plan(ctx, n);
    const int lanes = plan_lanes(plan);
    int splits = plan_splits(plan);
    if (splits == 0) splits = 1;                                   // a bare lanes value is a valid plan
    if (splits > ctx->prob.n_chunks && ctx->prob.n_chunks > 0)
        return fail(STN_ERR_INVALID, "plan asks for %d splits but the problem has %d chunks", splits,
                    ctx->prob.n_chunks);
    if (splits > 65535) return fail(STN_ERR_INVALID, "too many splits");
    double* partial = nullptr;
    if (splits > 1) {
        const int64_t need = (int64_t)splits * n;
        if (ctx->scratch_elems[scratch_slot] < need) {
            if (ctx->scratch[scratch_slot]) {
                // slot 0 may have been used from another of the caller's streams before
                if (scratch_slot == 0) STN_CUDA(cudaDeviceSynchronize());
                else STN_CUDA(cudaStreamSynchronize(st));
                STN_CUDA(cudaFree(ctx->scratch[scratch_slot]));
                ctx->scratch[scratch_slot] = nullptr;
                ctx->scratch_elems[scratch_slot] = 0;
            }
            STN_CUDA(cudaMalloc(&ctx->scratch[scratch_slot], (size_t)need * sizeof(double)));
            ctx->scratch_elems[scratch_slot] = need;
        }
        partial = ctx->scratch[scratch_slot];
    }
    STN_DBG_RANGE(theta, dbg_theta_bytes(n, ld, layout, ctx->prob.n_params), "theta");
    STN_DBG_RANGE(out, (size_t)n * sizeof(double), "out");
    if (partial) STN_DBG_RANGE(partial, (size_t)splits * n * sizeof(double), "split scratch");
    const int R = choose_R(ctx, n, lanes, splits);
    // EFAC-heavy problems use the R = STN_RBIG_EFAC kernel for large batches, others R = STN_RBIG
#define STN_GO(RR, LL) return launch_fast<RR, LL>(ctx, prob, theta, n, ld, layout, out, st, splits, partial)
    switch (lanes) {
        case 1:
            if (R > 1) {
                // Wave quantisation: every CTA of the big kernel runs for the same time, so with FEW waves (the
                // CTAs are still in lockstep) a last wave that is only partly full costs a whole wave: 4.6 waves
                // at 2^19 vectors take the time of 5.  The samples beyond the last FULL wave then go to the
                // R = 2 kernel in 64-thread CTAs: two thirds of the work per thread and many small CTAs that
                // spread evenly over the SMs (6.40 -> 6.14 ms at 2^19).  After ~10 waves the CTAs have drifted
                // apart and the partial wave costs little (23.4 ms at 2^21 without, 23.8 with): not used there.
                // R never changes a result bit, so this is invisible in the output.
                const int bigR = ctx->efac_heavy ? STN_RBIG_EFAC : STN_RBIG;
                const int64_t tile = (int64_t)threads_for(bigR) * bigR;
                const int64_t slots = (int64_t)ctx->sm_count * minblocks_for(bigR);
                const int64_t tiles = (n + tile - 1) / tile;
                const int64_t n_main = (tiles / slots) * slots * tile;
                constexpr int kTailR = STN_TAIL_R;
                const int64_t cap1 = (int64_t)ctx->sm_count * minblocks_tpb(kTailR, kTailThreads) * kTailThreads * kTailR;
                const int64_t left = tiles - (tiles / slots) * slots;                 // tiles of the last, partial wave
                const bool worth = tiles / slots <= 8 && 100 * left >= STN_TAIL_MIN_PCT * slots && 100 * left <= 85 * slots;
                if (splits == 1 && n_main > 0 && n_main < n && n - n_main <= cap1 && worth && STN_TAIL_SMALL_R) {
                    int rc = ctx->efac_heavy
                                 ? launch_fast<STN_RBIG_EFAC, 1>(ctx, prob, theta, n_main, ld, layout, out, st, 1, nullptr)
                                 : launch_fast<STN_RBIG, 1>(ctx, prob, theta, n_main, ld, layout, out, st, 1, nullptr);
                    if (rc) return rc;
                    ProbDev tail = prob;
                    for (int k = 0; k < tail.n_extra_out; ++k) tail.extra_out[k] += n_main;
                    const double* th_tail = layout == STN_LAYOUT_AOS ? theta + n_main * ld : theta + n_main;
                    return launch_fast<kTailR, 1, kTailThreads>(ctx, tail, th_tail, n - n_main, ld, layout, out + n_main, st, 1, nullptr);
                }
                if (ctx->efac_heavy) STN_GO(STN_RBIG_EFAC, 1);
                STN_GO(STN_RBIG, 1);
            }
            STN_GO(1, 1);
        case 2: STN_GO(1, 2);
        case 4: STN_GO(1, 4);
        case 8: STN_GO(1, 8);
        case 16: STN_GO(1, 16);
        case 32: STN_GO(1, 32);
        default: return fail(STN_ERR_INVALID, "plan lanes must be a power of two <= 32 (got %d)", lanes);
    }
#undef STN_GO
}

int check_call(StnCtx* ctx, const void* theta, int64_t n, int64_t ld, int layout, int mode,
               const void* out) {
    if (!ctx) return fail(STN_ERR_INVALID, "null context");
    if (n < 0) return fail(STN_ERR_INVALID, "negative batch size");
    if (n > 0 && (!theta || !out)) return fail(STN_ERR_INVALID, "null buffer");
    if (layout != STN_LAYOUT_AOS && layout != STN_LAYOUT_SOA) return fail(STN_ERR_INVALID, "bad layout %d", layout);
    if (mode != STN_MODE_FAST && mode != STN_MODE_STRICT) return fail(STN_ERR_INVALID, "bad mode %d", mode);
    if (layout == STN_LAYOUT_AOS && ld < ctx->prob.n_params)
        return fail(STN_ERR_INVALID, "AoS leading dimension %lld < P = %d", (long long)ld, ctx->prob.n_params);
    if (layout == STN_LAYOUT_SOA && ld < n)
        return fail(STN_ERR_INVALID, "SoA leading dimension %lld < n = %lld", (long long)ld, (long long)n);
    return STN_OK;
}


}  // namespace
