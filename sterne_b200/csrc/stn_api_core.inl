// C ABI, part 1: contexts, likelihood / posterior / chi-square / model launches, priors, host entry point,
// device memory and IPC plumbing, tooling.  Included by sterne_b200.cu.
extern "C" {

int stn_version(void) { return STN_ABI_VERSION; }

const char* stn_last_error(void) { return g_err.c_str(); }

int stn_create(const StnProblem* pb, int device, StnCtx** out_ctx) {
    if (!pb || !out_ctx) return fail(STN_ERR_INVALID, "null argument");
    *out_ctx = nullptr;
    if (pb->n_sources <= 0 || !pb->sources) return fail(STN_ERR_INVALID, "problem has no sources");
    if (pb->n_params <= 0) return fail(STN_ERR_INVALID, "problem has no parameters");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(STN_ERR_DEVICE, "no CUDA device visible (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(STN_ERR_INVALID, "device %d out of range (%d visible)", device, ndev);
    cudaDeviceProp prop;
    STN_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(STN_ERR_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    STN_CUDA(cudaSetDevice(device));

    StnCtx* ctx = new (std::nothrow) StnCtx();
    if (!ctx) return fail(STN_ERR_NOMEM, "out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;

    auto bail = [&](int code) {
        stn_destroy(ctx);
        return code;
    };

    std::vector<SrcDev> src(pb->n_sources);
    std::vector<ChunkDev> chunks;
    for (int s = 0; s < pb->n_sources; ++s) {
        const StnSource& hs = pb->sources[s];
        if (hs.n_epochs < 0 || hs.n_epochs > 0x3fffffff) return bail(fail(STN_ERR_INVALID, "source %d: bad epoch count", s));
        if (hs.n_epochs > 0 && (!hs.fast_rows || !hs.strict_rows))
            return bail(fail(STN_ERR_INVALID, "source %d: null epoch table", s));
        for (int r = 0; r < STN_NROOTS; ++r)
            if (hs.col[r] >= pb->n_params)
                return bail(fail(STN_ERR_INVALID, "source %d: column %d out of range (P = %d)", s, hs.col[r], pb->n_params));
        if (hs.efac_on && hs.col[STN_EFAC] < 0)
            return bail(fail(STN_ERR_INVALID, "source %d: efac_on without an efac column", s));
        SrcDev& d = src[s];
        std::memset(&d, 0, sizeof(d));
        int rc = upload(ctx, hs.fast_rows, (size_t)hs.n_epochs * STN_FAST_ROW, &d.fast_rows);
        if (rc) return bail(rc);
        rc = upload(ctx, hs.strict_rows, (size_t)hs.n_epochs * STN_STRICT_ROW, &d.strict_rows);
        if (rc) return bail(rc);
        d.n_epochs = (int)hs.n_epochs;
        for (int r = 0; r < STN_NROOTS; ++r) d.col[r] = hs.col[r];
        d.efac_on = hs.efac_on ? 1 : 0;
        d.binary = hs.binary ? 1 : 0;
        d.v_shift = 0;
        d.v_lo_bits = 0;
        d.v_hi = d.v_hib = 0.0;
        d.chi_scale = 1.0;
        if (d.efac_on && hs.n_epochs > 0) {
            // variance v = a + g*b with a = err_random^2, b = err_sys^2, g = efac^2 >= 0 (squared
            // radian-scale errors, ~1e-19).  The device copy of the four variance columns is stored
            // times 2^k with k chosen so that the scaled a's straddle 1: products of many variances
            // then stay in range.  The scaling is exact and undone exactly at the fold.
            double a_min = INFINITY, a_max = 0.0, b_max = 0.0;
            for (int64_t e = 0; e < hs.n_epochs; ++e) {
                const double* row = hs.fast_rows + e * STN_FAST_ROW;
                for (int k = 6; k < 8; ++k) { if (row[k] < a_min) a_min = row[k]; if (row[k] > a_max) a_max = row[k]; }
                for (int k = 8; k < 10; ++k) if (row[k] > b_max) b_max = row[k];
            }
            const bool usable = std::isfinite(a_max) && std::isfinite(b_max) && a_max > 0.0 && a_min >= 0.0;
            int k = 0;
            if (usable) {
                const double lo = a_min > 0.0 ? a_min : a_max;
                k = -(std::ilogb(lo) + std::ilogb(a_max)) / 2;
                if (k > 900) k = 900;
                if (k < -900) k = -900;
            }
            std::vector<double> scaled(hs.fast_rows, hs.fast_rows + (size_t)hs.n_epochs * STN_FAST_ROW);
            for (int64_t e = 0; e < hs.n_epochs; ++e)
                for (int c = 6; c < 10; ++c) scaled[(size_t)e * STN_FAST_ROW + c] = std::ldexp(scaled[(size_t)e * STN_FAST_ROW + c], k);
            if (cudaMemcpy(const_cast<double*>(d.fast_rows), scaled.data(), scaled.size() * sizeof(double),
                           cudaMemcpyHostToDevice) != cudaSuccess)
                return bail(fail(STN_ERR_CUDA, "source %d: upload of the scaled variance columns failed", s));
            d.v_shift = k;
            d.chi_scale = std::ldexp(1.0, k);
            d.v_hi = std::ldexp(a_max, k);
            d.v_hib = std::ldexp(b_max, k);
            const double lo_s = std::ldexp(a_min, k);
            // a == 0 somewhere (no random error): nothing bounds the variances from below, renormalise every epoch
            d.v_lo_bits = (usable && lo_s > 0.0 && std::isfinite(lo_s)) ? (lo_s >= 1.0 ? 0 : -std::ilogb(lo_s)) : 4096;
        }
        d.cos_decj = hs.binary ? hs.cos_decj : 1.0;
        d.inv_cos_decj = 1.0 / d.cos_decj;
        d.neg_sum_log_err = hs.neg_sum_log_err;
        d.a1_lts = hs.a1_lts;
        ctx->n_epochs.push_back(d.n_epochs);
        if (d.n_epochs > ctx->max_epochs) ctx->max_epochs = d.n_epochs;
        // chunk list: every source contributes at least one (possibly empty) chunk so that
        // its prologue / fold always runs
        int done = 0;
        do {
            ChunkDev c;
            c.rows = d.fast_rows + (size_t)done * STN_FAST_ROW;
            c.n_epochs = (d.n_epochs - done) < kChunk ? (d.n_epochs - done) : kChunk;
            c.source = s;
            c.flags = (done == 0 ? kFirst : 0);
            c.pad = 0;
            done += c.n_epochs;
            if (done >= d.n_epochs) c.flags |= kLast;
            chunks.push_back(c);
        } while (done < d.n_epochs);
    }
    {
        int64_t ef = 0, all = 0;
        for (const SrcDev& d : src) { all += d.n_epochs; if (d.efac_on) ef += d.n_epochs; }
        ctx->efac_heavy = all > 0 && 2 * ef >= all;
    }
    for (int k = 0; k < pb->n_a1dot; ++k)
        if (pb->a1dot_source[k] < 0 || pb->a1dot_source[k] >= pb->n_sources)
            return bail(fail(STN_ERR_INVALID, "a1dot constraint %d: bad source", k));
    for (int k = 0; k < pb->n_sini; ++k)
        if (pb->sini_col[k] < 0 || pb->sini_col[k] >= pb->n_params)
            return bail(fail(STN_ERR_INVALID, "sin(incl) constraint %d: bad column", k));

    ProbDev& P = ctx->prob;
    P.n_sources = pb->n_sources;
    P.n_params = pb->n_params;
    P.n_chunks = (int)chunks.size();
    P.n_a1dot = pb->n_a1dot;
    P.n_sini = pb->n_sini;
    P.mas2rad = pb->mas2rad > 0.0 ? pb->mas2rad : STN_MAS2RAD;
    P.deg2rad = pb->deg2rad > 0.0 ? pb->deg2rad : STN_DEG2RAD;
    P.masyr2rads = pb->masyr2rads > 0.0 ? pb->masyr2rads : STN_MASYR2RADS;
    int rc;
    if ((rc = upload(ctx, src.data(), src.size(), &P.src))) return bail(rc);
    if ((rc = upload(ctx, chunks.data(), chunks.size(), &P.chunks))) return bail(rc);
    {
        std::vector<int> first(pb->n_sources + 1, (int)chunks.size());
        for (int c = (int)chunks.size() - 1; c >= 0; --c) first[chunks[c].source] = c;
        if ((rc = upload(ctx, first.data(), first.size(), &P.first_chunk))) return bail(rc);
    }
    if ((rc = upload(ctx, pb->a1dot_source, (size_t)pb->n_a1dot, &P.a1_src))) return bail(rc);
    if ((rc = upload(ctx, pb->a1dot_mu, (size_t)pb->n_a1dot, &P.a1_mu))) return bail(rc);
    if ((rc = upload(ctx, pb->a1dot_sigma, (size_t)pb->n_a1dot, &P.a1_sigma))) return bail(rc);
    if ((rc = upload(ctx, pb->sini_col, (size_t)pb->n_sini, &P.sini_col))) return bail(rc);
    if ((rc = upload(ctx, pb->sini_mu, (size_t)pb->n_sini, &P.sini_mu))) return bail(rc);
    if ((rc = upload(ctx, pb->sini_sigma, (size_t)pb->n_sini, &P.sini_sigma))) return bail(rc);
    *out_ctx = ctx;
    return STN_OK;
}

int stn_destroy(StnCtx* ctx) {
    if (!ctx) return STN_OK;
    cudaSetDevice(ctx->device);
    for (void* p : ctx->dev_allocs) cudaFree(p);
    for (void* p : ctx->prior_allocs) cudaFree(p);
    for (int i = 0; i < 1 + StnCtx::kSlots; ++i)
        if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    if (ctx->small_in) cudaFreeHost(ctx->small_in);
    if (ctx->small_out) cudaFreeHost(ctx->small_out);
    for (int i = 0; i < StnCtx::kSlots; ++i) {
        if (ctx->stage_theta[i]) cudaFree(ctx->stage_theta[i]);
        if (ctx->stage_out[i]) cudaFree(ctx->stage_out[i]);
        if (ctx->pin_theta[i]) cudaFreeHost(ctx->pin_theta[i]);
        if (ctx->pin_out[i]) cudaFreeHost(ctx->pin_out[i]);
        if (ctx->h2d_done[i]) cudaEventDestroy(ctx->h2d_done[i]);
        if (ctx->d2h_done[i]) cudaEventDestroy(ctx->d2h_done[i]);
        if (ctx->streams[i]) cudaStreamDestroy(ctx->streams[i]);
    }
    if (ctx->gather_buf) cudaFree(ctx->gather_buf);
    if (ctx->sampler_buf) cudaFree(ctx->sampler_buf);
    if (ctx->sampler_nacc) cudaFree(ctx->sampler_nacc);
    for (int i = 0; i < 2; ++i)
        if (ctx->sampler_ev[i]) cudaEventDestroy(ctx->sampler_ev[i]);
    if (ctx->timing_stream) cudaStreamDestroy(ctx->timing_stream);
    delete ctx;
    return STN_OK;
}

int stn_auto_plan(StnCtx* ctx, int64_t n) {
    if (!ctx) return fail(STN_ERR_INVALID, "null context");
    return pick_plan(ctx, n);
}

int64_t stn_launch_count(StnCtx* ctx) { return ctx ? ctx->launches : 0; }

int stn_loglike(StnCtx* ctx, const double* theta_dev, int64_t n, int64_t ld, int layout, int mode,
                int plan, double* out_dev, void* stream) {
    int rc = check_call(ctx, theta_dev, n, ld, layout, mode, out_dev);
    if (rc) return rc;
    STN_CUDA(cudaSetDevice(ctx->device));
    return launch_loglike(ctx, theta_dev, n, ld, layout, mode, plan, out_dev, static_cast<cudaStream_t>(stream));
}

int stn_loglike_scatter(StnCtx* ctx, const double* theta_dev, int64_t n, int64_t ld, int layout, int mode,
                        int plan, double* const* outs, int n_outs, void* stream) {
    if (!outs || n_outs < 1 || n_outs > STN_MAX_OUTS) return fail(STN_ERR_INVALID, "1..%d output buffers are needed", STN_MAX_OUTS);
    for (int k = 0; k < n_outs; ++k)
        if (n > 0 && !outs[k]) return fail(STN_ERR_INVALID, "null output buffer %d", k);
    int rc = check_call(ctx, theta_dev, n, ld, layout, mode, outs[0]);
    if (rc) return rc;
    STN_CUDA(cudaSetDevice(ctx->device));
    return launch_loglike(ctx, theta_dev, n, ld, layout, mode, plan, outs[0], static_cast<cudaStream_t>(stream), 0, 0,
                          outs + 1, n_outs - 1);
}

int stn_dev_alloc(int device, void** ptr, int64_t bytes) {
    if (!ptr || bytes < 0) return fail(STN_ERR_INVALID, "bad argument");
    STN_CUDA(cudaSetDevice(device));
    STN_CUDA(cudaMalloc(ptr, (size_t)(bytes ? bytes : 1)));
    return STN_OK;
}

int stn_dev_free(int device, void* ptr) {
    if (!ptr) return STN_OK;
    STN_CUDA(cudaSetDevice(device));
    STN_CUDA(cudaFree(ptr));
    return STN_OK;
}

int stn_ipc_export(void* dev_ptr, void* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == STN_IPC_HANDLE_BYTES, "IPC handle size");
    if (!dev_ptr || !handle64) return fail(STN_ERR_INVALID, "null argument");
    cudaIpcMemHandle_t h;
    STN_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
    std::memcpy(handle64, &h, sizeof(h));
    return STN_OK;
}

int stn_ipc_open(int device, const void* handle64, void** peer_ptr) {
    if (!handle64 || !peer_ptr) return fail(STN_ERR_INVALID, "null argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, sizeof(h));
    STN_CUDA(cudaSetDevice(device));
    STN_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return STN_OK;
}

int stn_ipc_close(int device, void* peer_ptr) {
    if (!peer_ptr) return STN_OK;
    STN_CUDA(cudaSetDevice(device));
    STN_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return STN_OK;
}

int stn_set_prior(StnCtx* ctx, const StnPrior* pr) {
    if (!ctx || !pr) return fail(STN_ERR_INVALID, "null argument");
    if (pr->n_params != ctx->prob.n_params)
        return fail(STN_ERR_INVALID, "prior has %d entries, the problem has %d parameters", pr->n_params,
                    ctx->prob.n_params);
    if (!pr->kind || !pr->a || !pr->b) return fail(STN_ERR_INVALID, "null prior arrays");
    double c = 0.0;
    for (int k = 0; k < pr->n_params; ++k) {
        const double a = pr->a[k], b = pr->b[k];
        switch (pr->kind[k]) {
            case STN_PRIOR_UNIFORM:
                if (!(b > a)) return fail(STN_ERR_INVALID, "prior %d: Uniform needs a < b", k);
                c += -std::log(b - a);
                break;
            case STN_PRIOR_GAUSSIAN:
                if (!(b > 0)) return fail(STN_ERR_INVALID, "prior %d: Gaussian needs sigma > 0", k);
                c += -std::log(b) - 0.5 * std::log(2.0 * 3.14159265358979323846);
                break;
            case STN_PRIOR_SINE:
                if (!(b > a) || a < 0.0 || b > 3.14159265358979323846)
                    return fail(STN_ERR_INVALID, "prior %d: Sine needs 0 <= a < b <= pi", k);
                c += -std::log(std::cos(a) - std::cos(b));
                break;
            default: return fail(STN_ERR_INVALID, "prior %d: unknown kind %d", k, pr->kind[k]);
        }
    }
    for (int k = 0; k < pr->n_sin_limits; ++k)
        if (pr->sin_col[k] < 0 || pr->sin_col[k] >= pr->n_params)
            return fail(STN_ERR_INVALID, "sin limit %d: bad column", k);
    STN_CUDA(cudaSetDevice(ctx->device));
    STN_CUDA(cudaDeviceSynchronize());                 // no launch may still read the old prior
    // only the previous prior's arrays are replaced; scratch / staging buffers stay as they are
    for (void* p : ctx->prior_allocs) cudaFree(p);
    ctx->prior_allocs.clear();
    ctx->has_prior = false;
    ProbDev P = ctx->prob;
    // upload() records into dev_allocs; move those entries to prior_allocs so they can be replaced
    const size_t mark = ctx->dev_allocs.size();
    int rc = STN_OK;
    if (!rc) rc = upload(ctx, pr->kind, (size_t)pr->n_params, &P.prior_kind);
    if (!rc) rc = upload(ctx, pr->a, (size_t)pr->n_params, &P.prior_a);
    if (!rc) rc = upload(ctx, pr->b, (size_t)pr->n_params, &P.prior_b);
    if (!rc) rc = upload(ctx, pr->sin_col, (size_t)pr->n_sin_limits, &P.sinlim_col);
    if (!rc) rc = upload(ctx, pr->sin_lo, (size_t)pr->n_sin_limits, &P.sinlim_lo);
    if (!rc) rc = upload(ctx, pr->sin_hi, (size_t)pr->n_sin_limits, &P.sinlim_hi);
    ctx->prior_allocs.assign(ctx->dev_allocs.begin() + mark, ctx->dev_allocs.end());
    ctx->dev_allocs.resize(mark);
    if (rc) return rc;
    P.n_prior = pr->n_params;
    P.n_sinlim = pr->n_sin_limits;
    P.prior_const = c;
    ctx->prior_prob = P;
    ctx->has_prior = true;
    return STN_OK;
}

int stn_logpost(StnCtx* ctx, const double* theta_dev, int64_t n, int64_t ld, int layout, int mode,
                int plan, double* out_dev, void* stream) {
    int rc = check_call(ctx, theta_dev, n, ld, layout, mode, out_dev);
    if (rc) return rc;
    if (!ctx->has_prior) return fail(STN_ERR_INVALID, "stn_logpost needs stn_set_prior first");
    STN_CUDA(cudaSetDevice(ctx->device));
    return launch_loglike(ctx, theta_dev, n, ld, layout, mode, plan, out_dev, static_cast<cudaStream_t>(stream),
                          1);
}

int stn_chisq(StnCtx* ctx, const double* theta_dev, int64_t n, int64_t ld, int layout, int mode,
              int plan, double* out_dev, void* stream) {
    int rc = check_call(ctx, theta_dev, n, ld, layout, mode, out_dev);
    if (rc) return rc;
    STN_CUDA(cudaSetDevice(ctx->device));
    return launch_loglike(ctx, theta_dev, n, ld, layout, mode, plan, out_dev, static_cast<cudaStream_t>(stream),
                          2);
}

int stn_loglike_timed(StnCtx* ctx, const double* theta_dev, int64_t n, int64_t ld, int layout,
                      int mode, int plan, double* out_dev, int steps, double* total_ms) {
    int rc = check_call(ctx, theta_dev, n, ld, layout, mode, out_dev);
    if (rc) return rc;
    if (steps <= 0 || !total_ms) return fail(STN_ERR_INVALID, "bad steps / null total_ms");
    STN_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->timing_stream) STN_CUDA(cudaStreamCreateWithFlags(&ctx->timing_stream, cudaStreamNonBlocking));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float ms = 0.f;
    auto run = [&]() -> int {
        STN_CUDA(cudaEventCreate(&e0));
        STN_CUDA(cudaEventCreate(&e1));
        STN_CUDA(cudaDeviceSynchronize());
        STN_CUDA(cudaEventRecord(e0, ctx->timing_stream));
        for (int i = 0; i < steps; ++i) {
            const int r = launch_loglike(ctx, theta_dev, n, ld, layout, mode, plan, out_dev, ctx->timing_stream);
            if (r) return r;
        }
        STN_CUDA(cudaEventRecord(e1, ctx->timing_stream));
        STN_CUDA(cudaEventSynchronize(e1));
        STN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        return STN_OK;
    };
    rc = run();
    if (rc) cudaStreamSynchronize(ctx->timing_stream);    // nothing of this call may still be running on return
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (rc) return rc;
    *total_ms = ms;
    return STN_OK;
}

int stn_loglike_host(StnCtx* ctx, const double* theta_host, int64_t n, int64_t ld, int layout,
                     int mode, int plan, double* out_host) {
    int rc = check_call(ctx, theta_host, n, ld, layout, mode, out_host);
    if (rc) return rc;
    if (n == 0) return STN_OK;
    STN_CUDA(cudaSetDevice(ctx->device));
    if (mode == STN_MODE_FAST && plan == 0) plan = pick_plan(ctx, n);    // from the whole batch
    return run_host(ctx, theta_host, n, ld, layout, mode, plan, out_host);
}

int stn_host_hint(StnCtx* ctx, int concurrent_devices) {
    if (!ctx || concurrent_devices < 1) return fail(STN_ERR_INVALID, "bad argument");
    ctx->host_share = concurrent_devices;
    return STN_OK;
}

int stn_host_register(void* ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return fail(STN_ERR_INVALID, "bad argument");
    STN_CUDA(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
    return STN_OK;
}

int stn_host_unregister(void* ptr) {
    if (ptr) STN_CUDA(cudaHostUnregister(ptr));
    return STN_OK;
}

int stn_model(StnCtx* ctx, int source, const double* theta_dev, int64_t n, int64_t ld, int layout,
              int mode, double* radec_dev, double* off_mas_dev, void* stream) {
    int rc = check_call(ctx, theta_dev, n, ld, layout, mode, radec_dev);
    if (rc) return rc;
    if (source < 0 || source >= ctx->prob.n_sources) return fail(STN_ERR_INVALID, "source %d out of range", source);
    const int64_t total = n * (int64_t)ctx->n_epochs[source];
    if (total == 0) return STN_OK;
    STN_CUDA(cudaSetDevice(ctx->device));
    const int64_t grid = (total + kThreads - 1) / kThreads;
    if (grid > 0x7fffffffLL) return fail(STN_ERR_INVALID, "batch too large for one launch");
    STN_DBG_RANGE(theta_dev, dbg_theta_bytes(n, ld, layout, ctx->prob.n_params), "theta");
    STN_DBG_RANGE(radec_dev, (size_t)total * 2 * sizeof(double), "radec");
    if (off_mas_dev) STN_DBG_RANGE(off_mas_dev, (size_t)total * 2 * sizeof(double), "off_mas");
    (void)cudaGetLastError();
    stn_model_kernel<<<(unsigned)grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        ctx->prob, source, theta_dev, n, ld, layout, mode, radec_dev, off_mas_dev);
    STN_CUDA(cudaGetLastError());
    ctx->launches++;
    return STN_OK;
}

int stn_host_alloc(void** ptr, int64_t bytes) {
    if (!ptr || bytes < 0) return fail(STN_ERR_INVALID, "bad argument");
    STN_CUDA(cudaHostAlloc(ptr, (size_t)(bytes ? bytes : 1), cudaHostAllocPortable));
    return STN_OK;
}

int stn_host_free(void* ptr) {
    if (ptr) STN_CUDA(cudaFreeHost(ptr));
    return STN_OK;
}

int stn_fp64_peak(int device, int iters, int reps, double* tflops, double* ms_out) {
    if (iters <= 0 || reps <= 0 || !tflops) return fail(STN_ERR_INVALID, "bad argument");
    STN_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    STN_CUDA(cudaGetDeviceProperties(&prop, device));
    double* sink = nullptr;
    STN_CUDA(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    STN_CUDA(cudaEventCreate(&e0));
    STN_CUDA(cudaEventCreate(&e1));
    const int grid = prop.multiProcessorCount * 8;
    const int block = 256;
    double best = 1e300;
    for (int r = 0; r < reps + 1; ++r) {   // first launch is the warm-up
        STN_CUDA(cudaEventRecord(e0, 0));
        stn_dfma_peak_kernel<<<grid, block>>>(sink, iters, 0.999999, 1e-9);
        STN_CUDA(cudaGetLastError());
        STN_CUDA(cudaEventRecord(e1, 0));
        STN_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        STN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    const double flops = 2.0 * 8.0 * (double)iters * (double)grid * (double)block;
    *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return STN_OK;
}

}  // extern "C"
